"""GAN task base placeholder (reference generative/gan.py:12-149).  The two-loss simultaneous
D/G update is lowered in plan.py once enabled; until then constructing a GAN raises clearly."""
from .convnet import ConvNet


class GAN(ConvNet):
    def _init_model(self, **kwargs):
        raise NotImplementedError('GAN training (two losses, D/G variable split, generative/gan.py) '
                                  'is not lowered yet on the B200 backend')
