"""PixelMetric (train.py:17, eval.py:48): confusion matrix over the valid pixels -> per-class IoU / F1, OA, mIoU"""
import os

import numpy as np
import torch


class PixelMetric(object):
    def __init__(self, num_classes, logdir=None, logger=None):
        self.num_classes, self.logdir, self.logger = int(num_classes), logdir, logger
        self.cm = np.zeros((self.num_classes, self.num_classes), dtype=np.int64)      # rows = truth, columns = prediction

    def forward(self, y_true, y_pred):
        t = torch.as_tensor(y_true).reshape(-1).to(torch.int64)
        p = torch.as_tensor(y_pred).reshape(-1).to(torch.int64)
        k = self.num_classes
        self.cm += torch.bincount(t * k + p, minlength=k * k).reshape(k, k).cpu().numpy()

    def summary_all(self):
        cm = self.cm.astype(np.float64)
        tp = np.diag(cm)
        iou = tp / np.maximum(cm.sum(0) + cm.sum(1) - tp, 1)
        f1 = 2 * tp / np.maximum(cm.sum(0) + cm.sum(1), 1)
        out = dict(iou=iou, f1=f1, miou=float(iou.mean()), oa=float(tp.sum() / max(cm.sum(), 1)))
        msg = "mIoU %.4f  OA %.4f  IoU %s" % (out["miou"], out["oa"], np.array2string(iou, precision=4))
        (self.logger.info if self.logger is not None else print)(msg)
        if self.logdir:
            os.makedirs(self.logdir, exist_ok=True)
            np.save(os.path.join(self.logdir, "confusion_matrix.npy"), self.cm)
        return out
