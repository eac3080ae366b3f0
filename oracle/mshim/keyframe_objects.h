// Test infrastructure, never shipped.  Force-included ahead of the reference's src/KeyFrame.cc, src/Frame.cc and src/ORBmatcher.cc
// (oracle/Makefile refkeyframe): keeps the reference's own include/KeyFrame.h, include/Frame.h and include/ORBmatcher.h and
// replaces -- through their include guards -- MapPoint.h (stand-in), Map.h and KeyFrameDatabase.h (empty stubs) next to the
// extractor / vocabulary / converter stubs of frame_objects.h.
#pragma once
#define MSHIM_REAL_KEYFRAME
#define MAP_H
#define KEYFRAMEDATABASE_H
#include "frame_objects.h"

namespace ORB_SLAM2 {
class Map {
public:
    void EraseKeyFrame(KeyFrame*) {}
};
class KeyFrameDatabase {
public:
    void erase(KeyFrame*) {}
};
}  // namespace ORB_SLAM2
