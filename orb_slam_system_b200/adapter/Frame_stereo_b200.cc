// Frame::ComputeStereoMatches over liborb_b200: the one member function of the reference's src/Frame.cc (:446-619) that
// is part of the hot path.  Delete the function body from src/Frame.cc (or build Frame.cc with the symbol weakened, as
// oracle/Makefile adapterframe does) and add this file: the stereo constructor (src/Frame.cc:41-97) then calls it as
// before.  Row-band Hamming search, 11 x 11 SAD refinement, parabola fit, disparity gate and the median cut all run
// in orb_compute_stereo_matches; the pyramids are read where the two ORB_SLAM2::ORBextractor instances
// (adapter/ORBextractor.h) left them on the device, so mvImagePyramid is not touched.
#include "Frame.h"

#include <cstdlib>

#include "orb_match_b200.h"

namespace ORB_SLAM2 {

void Frame::ComputeStereoMatches() {
    // one GPU matcher handle per calling thread (the tracking thread builds the frames)
    thread_local orb_b200::Matcher gpu(getenv("ORB_B200_DEVICE") ? atoi(getenv("ORB_B200_DEVICE")) : 0);
    // mvuRight / mvDepth = N x -1.0f, then the matches (:448-449, :600-601, :614-617).  The disparity range comes from the
    // member `mb` exactly as in the reference (minZ = mb, :476-478): the constructor calls this function before it assigns the
    // static fx and `mb = mbf / fx` (:70 vs :86-94), so deriving mb from fx here would read a value that is not set yet.
    gpu.ComputeStereoMatchesMb(mpORBextractorLeft->handle(), mpORBextractorRight->handle(), mvKeys, mDescriptors, mvKeysRight,
                               mDescriptorsRight, mbf, mb, mvuRight, mvDepth);
}

}  // namespace ORB_SLAM2
