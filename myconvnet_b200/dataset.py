"""In-memory data set with the reference's DataSet surface.

Mirrors `DataSet(image_dirs, label_dirs, ..., from_memory=True, **kwargs)` of the reference
(dataset.py:34-147) for the path the training step needs: images and labels are numpy arrays held
in host memory, every global batch of `batch_size` examples is split into `num_shards` contiguous
per-device shards (dataset.py:104-113: `np.split(batch, num_shards)`), the remainder that does not
fill a global batch is dropped from training steps (optimizers.py:411-414).  One process per GPU
replaces the reference's per-device tf.data iterators: `shard_batches(rank)` yields this rank's
shard of every step.  Decoding image FILES (from_memory=False: cv2 / tf.py_func loaders, resize
and augmentation, dataset.py:134-138 and convnet.py:714-1135) is outside the hot path and raises.

Images may be float32 in [0,1] (what the reference's loaders produce) or raw uint8 — the network's
device prologue (mcn_input_prep) divides uint8 by 255, subtracts `image_mean`, centre-crops to the
network input and scales (convnet.py:449-471), so uint8 batches cross PCIe at a quarter of the bytes.
"""
import numpy as np


class DataSet(object):
    IMAGE_ONLY = 'image_only'
    IMAGE_CLASSIFICATION = 'image_classification'
    IMAGE_SEGMENTATION = 'image_segmentation'
    DCGAN = 'dcgan'

    def __init__(self, image_dirs, label_dirs=None, task_type=IMAGE_ONLY, class_names=None, num_classes=None,
                 out_size=None, resize_method=None, resize_randomness=False, shuffle_data=None,
                 from_memory=False, **kwargs):
        if not from_memory:
            raise NotImplementedError('DataSet(from_memory=False) reads and decodes image files; only the '
                                      'in-memory path (numpy arrays) belongs to the training-step hot path')
        if image_dirs is None:
            raise ValueError('from_memory=True needs the image array')
        self._images = np.ascontiguousarray(image_dirs)
        if self._images.dtype not in (np.uint8, np.float32):
            self._images = self._images.astype(np.float32)
        n = len(self._images)
        if label_dirs is None:
            label_dirs = np.full((n,), np.nan, dtype=np.float32)        # fake labels (dataset.py:50-53)
        self._labels = np.ascontiguousarray(label_dirs)
        assert len(self._labels) == n, 'Number of examples mismatch, between images and labels'
        self._image_size = tuple(out_size) if out_size is not None else tuple(self._images.shape[1:])
        self._task_type = task_type
        self._shuffle = kwargs.get('shuffle', True) if shuffle_data is None else shuffle_data
        self._from_memory = True
        if class_names is None:
            if task_type in (DataSet.IMAGE_CLASSIFICATION, DataSet.IMAGE_SEGMENTATION, DataSet.DCGAN):
                assert num_classes is not None, 'Either class_names or num_classes must be provided.'
            self._num_classes = num_classes
        else:
            self._num_classes = len(class_names)
        self._class_names = class_names
        self._num_shards = max(1, int(kwargs.get('num_gpus', 1) or 1))
        self._batch_size = int(kwargs.get('batch_size', 16))
        self._image_mean = kwargs.get('image_mean', 0.5)
        self._parameters = kwargs
        self._rng = np.random.default_rng(kwargs.get('shuffle_seed', 0))

    # ---- the reference's read-only properties (dataset.py:149-230)
    image_dirs = property(lambda self: self._images)
    label_dirs = property(lambda self: self._labels)
    image_size = property(lambda self: self._image_size)
    task_type = property(lambda self: self._task_type)
    shuffle = property(lambda self: self._shuffle)
    from_memory = property(lambda self: self._from_memory)
    num_classes = property(lambda self: self._num_classes)
    class_names = property(lambda self: self._class_names)
    num_shards = property(lambda self: self._num_shards)
    batch_size = property(lambda self: self._batch_size)
    image_mean = property(lambda self: self._image_mean)
    num_examples = property(lambda self: len(self._images))

    def __len__(self):
        return self.num_examples

    @property
    def input_dtype(self):
        """'u8' or 'f32': what ConvNet(input_dtype=...) must be built with."""
        return 'u8' if self._images.dtype == np.uint8 else 'f32'

    def labels_for_device(self):
        """Labels as the device takes them: int32 class indices, NaN (fake label) -> -1
        (convnet.py:441-449: one_hot(int(-1)) is the all-zero row)."""
        lab = self._labels
        if np.issubdtype(lab.dtype, np.floating):
            lab = np.where(np.isnan(lab), -1.0, lab)
        return lab.astype(np.int32)

    def shard_batches(self, rank=0, epoch_seed=None):
        """Yields (images, labels) of this rank's shard for every full global batch of one epoch
        (global batch = batch_size, shard = batch_size // num_shards contiguous examples)."""
        n = self.num_examples
        order = (np.random.default_rng(epoch_seed) if epoch_seed is not None else self._rng).permutation(n) \
            if self._shuffle else np.arange(n)
        per = self._batch_size // self._num_shards
        labels = self.labels_for_device()
        for s in range(n // self._batch_size):
            idx = order[s * self._batch_size:(s + 1) * self._batch_size][rank * per:(rank + 1) * per]
            yield self._images[idx], labels[idx]
