"""Extracts the 4 calibrated demo cameras of the reference (assets/demo/{R_list,t_list,intr_list}.npy,
camera-to-world, OpenCV convention; consumed at /root/reference/src/demo.py:125-135) into a small JSON
fixture so that bench.py / tests never read /root/reference at run time.  Run in the build container only."""
import json, os, sys
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
d = os.path.join(ref, "assets", "demo")
R = np.load(os.path.join(d, "R_list.npy")); t = np.load(os.path.join(d, "t_list.npy")); K = np.load(os.path.join(d, "intr_list.npy"))
cams = []
for i in range(R.shape[0]):
    c2w = np.eye(4); c2w[:3, :3] = R[i]; c2w[:3, 3] = t[i]
    w2c = np.linalg.inv(c2w)   # /root/reference/src/real_world/gs/trainer.py:15-18
    cams.append(dict(k=K[i].tolist(), w2c=w2c.tolist()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gs_dynamics_b200", "data", "demo_cameras.json")
json.dump(dict(w=640, h=480, source="robo-alex/gs-dynamics assets/demo R_list/t_list/intr_list.npy", cams=cams), open(out, "w"), indent=1)
print("wrote", out)
