import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from fusion_power_video_b200 import synth, host
W,H=1280,800
fr=synth.plasma_frames(256,W,H,bits=12,seed=1).reshape(256,-1)
fr=np.ascontiguousarray(np.tile(fr,(4,1)))
for ct in ("5",):
    os.environ["FPV_COPY_THREADS"]=ct
    for batch in (16,32,64,128):
        host.time_encode(fr[:64],W,H,4,False,threads=16,batch=batch,gpu_entropy=True)
        best=min(host.time_encode(fr,W,H,4,False,threads=16,batch=batch,gpu_entropy=True)[0] for _ in range(3))
        print("copy threads",ct,"batch",batch,"GB/s",round(fr.size*2/best/1e9,2),"fps",round(fr.shape[0]/best))
