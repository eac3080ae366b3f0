// Stand-in so the reference sources compile without OpenCV; see ../cvshim.hpp
#include "../../cvshim.hpp"
