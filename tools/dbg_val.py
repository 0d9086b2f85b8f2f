import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yolo11_ref as R
from yolo_infer_b200.engine import YOLO
from yolo_infer_b200 import val as V
ref = R.build("n", init="calibrated", seed=0)
eng = YOLO.from_state_dict({k: v.clone() for k, v in ref.state_dict().items()}, "n").to("cuda:0")
rng = np.random.default_rng(5)
img = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
det = eng.predict(img, conf=0.4, iou=0.6, multi_label=True, verbose=False)[0].cpu().boxes.data.numpy()
allp = eng.predict(img, conf=0.001, iou=0.6, multi_label=True, verbose=False)[0].cpu().boxes.data.numpy()
print('labels', det.shape, 'preds', allp.shape, 'pred conf range', allp[:,4].min(), allp[:,4].max())
print('n preds above .4:', (allp[:,4] > 0.4).sum())
gt = np.concatenate((det[:, 5:6], det[:, :4]), 1)
m = V.evaluate([allp], [gt], 80)
print('map50', m.map50, 'mp', m.mp, 'mr', m.mr, 'classes', len(m.ap_class_index))
tp = V.match_predictions(allp[:,5], gt[:,0], V.box_iou(gt[:,1:], allp[:,:4]))
print('tp@.5 among top', tp[:len(det),0].mean(), 'total tp', tp[:,0].sum())
# are the label boxes present in allp?
for r in det[:5]:
    d = np.abs(allp[:, :4] - r[:4]).max(1)
    j = d.argmin(); print(r, '->', allp[j], d[j])
