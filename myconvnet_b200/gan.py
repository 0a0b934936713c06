"""GAN task base on the graph facade (semantics of reference generative/gan.py:12-149 and the
variable split / simultaneous update of generative/optimizers_gan.py:17-121).

Per tower the reference builds D(real) -> G(z) -> D(G(z)) with shared discriminator variables
(gan.py:61-65) and two losses (gan.py:125-138):
    loss_d = mean(w[1]*sigCE(1, D(x)) + w[0]*sigCE(0, D(G(z)))),  loss_g = mean(w[0]*sigCE(1, D(G(z))))
with one-sided label smoothing labels*(1-ls) (gan.py:145-149) and NO L2 term.  The optimiser takes
d(loss_d)/d(theta_D) and d(generator_scaling_factor*loss_g)/d(theta_G) from the same forward and
applies both in one update.  plan.py lowers this as two backward passes over the shared graph.
`Y` is the latent batch [N, num_classes] (num_classes = latent length, gan.py:35); the reference
replaces NaN rows by U(-1,1) noise (gan.py:37) — here the caller supplies the noise.
D batch-norm moving statistics see the real batch first and the generated batch second (the
reference leaves the order of its two assigns undefined).
"""
import numpy as np

from .convnet import ConvNet


class GAN(ConvNet):
    uses_l2 = False

    @property
    def num_blocks_g(self):
        return self._num_blocks_g

    def _init_model(self, **kwargs):
        self._curr_device = 0
        self._curr_block = None
        self._num_blocks_g = 1
        self.X_in, self.X = self._make_inputs()
        n = self._batch_size
        self.Y_in = self.graph.placeholder('Y', (n, self.num_classes), 'f32')
        self.Y = self.Y_in
        if self.dtype != 'f32':
            self.Y = self.graph._add('cast', [self.Y_in], [self.Y_in.shape], [self.dtype]).outputs[0]
        d_real = self._build_model()
        self.d = self._build_model_g()
        self.X = self.d['generate']
        self.generate = self.d['generate']
        self._reuse = True
        d_fake = self._build_model()
        self.logits_real = self._to_f32(d_real['logits'])
        self.logits_fake = self._to_f32(d_fake['logits'])
        self.d_real = d_real
        self.d.update(d_fake)
        self.d['logits_real'] = self.logits_real
        self.d['logits_fake'] = self.logits_fake
        self.dicts.append(self.d)
        self.pred = self.d['generate']
        self.losses_g = []
        loss_d, loss_g = self._build_loss(**kwargs)
        self.losses.append(loss_d)
        self.losses_g.append(loss_g)
        self.loss, self.loss_g = loss_d, loss_g
        blocks = sorted(b for b in self._block_list if b is not None)
        self._gan_node.attrs['g_blocks'] = set(blocks[len(blocks) - self.num_blocks_g:])

    def _build_model_g(self):
        raise NotImplementedError

    def _build_loss(self, **kwargs):
        w = self.loss_weights
        w = np.ones(2, dtype=np.float32) if w is None else np.array(w, dtype=np.float32)
        node = self.graph._add('gan_loss', [self.logits_real, self.logits_fake], [(), ()], ['f32', 'f32'],
                               {'w': (float(w[0]), float(w[1])),
                                'label_smoothing': float(kwargs.get('label_smoothing', 0.0)),
                                'generator_scaling_factor': float(kwargs.get('generator_scaling_factor', 1.0)),
                                'rows': int(np.prod(self.logits_real.shape)),
                                'generate': self.generate, 'g_blocks': None})
        self._gan_node = node
        self.graph.losses.extend(node.outputs)
        return node.outputs[0], node.outputs[1]

    def gan_variable_split(self):
        """(D variables, G variables) by block (optimizers_gan.py:22-30): the last num_blocks_g
        integer blocks are the generator, the others and block None the discriminator."""
        blocks = sorted(b for b in self.block_list if b is not None)
        g_blocks = set(blocks[len(blocks) - self.num_blocks_g:])
        vd = [v for v in self.graph.vars.values() if v.block not in g_blocks]
        vg = [v for v in self.graph.vars.values() if v.block in g_blocks]
        return vd, vg
