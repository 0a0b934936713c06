"""ORACLE (test infrastructure): eager SegNet base restating reference segmentation/segnet.py:11-121
for one tower: labels 0 = ignore, 1..C = classes -> one-hot of round(Y-1) (zero row for ignored
pixels), backbone built with backbone_only then _build_model_seg, per-pixel softmax CE averaged
over all pixels."""
import numpy as np
import torch

from myconvnet_b200 import tfshim as tf
from . import tf_ops as ops
from .ref_convnet import ConvNet, OTensor


class SegNet(ConvNet):
    _spatial_label_smoothing = True      # segnet.py:116-121: 5x5 SAME average of the one-hot map

    def forward(self, X, Y):
        tf.reset_scopes()
        from . import ref_convnet
        ref_convnet._CURRENT[0] = self
        self._random_layers = 0
        self._relu_calls = 0
        self._block_list = []
        self.collections = {}
        self.bn_updates = {}
        for name, t in self.vars.items():
            t.requires_grad_(self.var_meta.get(name, {}).get("trainable", True)
                             and self.var_meta.get(name, {}).get("kind", "weight") != "stat")
            t.grad = None
        x = torch.as_tensor(np.asarray(X), dtype=self.dtype)
        self.X = OTensor(self._q((x - self.image_mean) * self.scale_factor))
        self.Y = torch.as_tensor(np.asarray(Y)).long() - 1          # -1 = ignore
        self._curr_block = None
        self._backbone_only = True
        d_backbone = self._build_model()
        self._backbone_only = False
        self.d = self._build_model_seg(d_backbone)
        self.d.update(d_backbone)
        self.logits = self.d["logits"].t.to(self.dtype)
        self.pred = self.d["pred"].t
        return self._build_loss(**self._parameters)
