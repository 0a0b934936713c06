"""SegNet task base on the graph facade (semantics of reference segmentation/segnet.py:11-121).

Labels arrive as in the reference: float/int [N,H,W] with 0 = ignore and 1..C = classes
(README.md:48); the reference turns them into one-hot of round(Y-1), so ignored pixels are zero
rows (segnet.py:50-51) that contribute nothing to the loss.  On device the label tensor holds the
class index (Y-1, -1 = ignore): `label_offset = 1` tells the engine to subtract it while copying.
The loss is the inherited per-pixel softmax CE averaged over ALL pixels (convnet.py:594).
"""
from .convnet import ConvNet


class SegNet(ConvNet):
    label_offset = 1

    def _init_model(self, **kwargs):
        self._curr_device = 0
        self._curr_block = None
        self.X_in, self.X = self._make_inputs()
        n = self._batch_size
        h, w, _ = self.input_size
        self.Y = self.graph.placeholder('Y', (n, h, w), 'i32')
        self._backbone_only = True
        d_backbone = self._build_model()
        self._backbone_only = False
        self.d = self._build_model_seg(d_backbone)
        self._reuse = True
        self.logits = self._to_f32(self.d['logits'])
        self.d['logits'] = self.logits
        self.pred = self.d['pred']
        self.d.update(d_backbone)
        self.dicts.append(self.d)
        self.losses.append(self._build_loss(**kwargs))
        self.loss = self.losses[0]

    def _build_model(self):
        raise NotImplementedError

    def _build_model_seg(self, d_backbone):
        raise NotImplementedError

    def _label_smoothing_map(self):
        # labels*(1-ls) + ls*avg_pool2d(labels, 5x5, SAME) (segnet.py:116-121): evaluated inside the
        # loss kernel from the integer label map (mcn_softmax_xent seg_h / seg_w)
        return (int(self.input_size[0]), int(self.input_size[1]))
