"""ORACLE (test infrastructure): eager GAN base restating reference generative/gan.py:17-149 and
the D/G variable split of generative/optimizers_gan.py:22-30 for one tower."""
import numpy as np
import torch

from myconvnet_b200 import tfshim as tf
from . import tf_ops as ops
from .ref_convnet import ConvNet, OTensor


class GAN(ConvNet):
    @property
    def num_blocks_g(self):
        return self._num_blocks_g

    def forward(self, X, Z):
        """X real images in [0,1], Z latent vectors.  Returns (loss_d, loss_g)."""
        tf.reset_scopes()
        from . import ref_convnet
        ref_convnet._CURRENT[0] = self
        self._random_layers = 0
        self._relu_calls = 0
        self._block_list = []
        self.collections = {}
        self.bn_updates = {}
        self._num_blocks_g = 1
        for name, t in self.vars.items():
            t.requires_grad_(self.var_meta.get(name, {}).get("kind", "weight") != "stat")
            t.grad = None
        x = torch.as_tensor(np.asarray(X), dtype=self.dtype)
        self.X = OTensor(self._q((x - self.image_mean) * self.scale_factor))
        self.Y = OTensor(self._q(torch.as_tensor(np.asarray(Z), dtype=self.dtype)))
        self._curr_block = None
        d_real = self._build_model()
        self.d = self._build_model_g()
        self.X = self.d["generate"]
        d_fake = self._build_model()
        self.d.update(d_fake)
        self.logits_real = d_real["logits"].t.to(self.dtype)
        self.logits_fake = d_fake["logits"].t.to(self.dtype)
        w = self.loss_weights
        w = np.ones(2, dtype=np.float32) if w is None else np.array(w, dtype=np.float32)
        ls = self._parameters.get("label_smoothing", 0.0)
        ones = torch.ones_like(self.logits_real) * (1.0 - ls)         # one-sided smoothing
        zeros = torch.zeros_like(self.logits_fake)
        l_real = ops.sigmoid_cross_entropy(self.logits_real, ones)
        l_fake = ops.sigmoid_cross_entropy(self.logits_fake, zeros)
        l_g = ops.sigmoid_cross_entropy(self.logits_fake, ones)
        self.loss_d = (float(w[1]) * l_real + float(w[0]) * l_fake).mean()
        self.loss_g = (float(w[0]) * l_g).mean()
        return self.loss_d, self.loss_g

    def variable_split(self):
        blocks = sorted(b for b in self._block_list if b is not None)
        g_blocks = set(blocks[len(blocks) - self.num_blocks_g:])
        vd = [k for k, m in self.var_meta.items() if m["block"] not in g_blocks]
        vg = [k for k, m in self.var_meta.items() if m["block"] in g_blocks]
        return vd, vg
