#include "../mshim_cv.hpp"
