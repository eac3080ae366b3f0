import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from orb_slam_system_b200 import ORBextractor
from orb_slam_system_b200.synth import synth_frame
img=synth_frame(480,640)
ex=ORBextractor(1000,1.2,8,20,7)
try:
    k,d=ex(img); print(len(k))
except Exception as e:
    print('ERR',e)
