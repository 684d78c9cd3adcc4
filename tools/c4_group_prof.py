import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiducials_b200 import synth
from fiducials_b200.node import Detector, default_params
W, H, n, d = synth.CONFIGS["C4"]
bgr = synth.make_config_frame("C4", 0)[0]
det = Detector(default_params(dictionary=d), 0, W, H, 1)
for _ in range(3):
    ids, c = det.detect(bgr)
print(len(ids), {k: round(v, 3) for k, v in det.last_stage_ms().items() if v > 0.01})
