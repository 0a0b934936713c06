"""Import the reference's model files UNCHANGED against a ConvNet facade.

The reference model files (models/resnet_v1_5.py, efficientnet.py, deeplabv3plus.py, dcgan.py,
resnet_v1_5_dilated.py) start with ``import tensorflow.compat.v1 as tf`` and
``from convnet import ConvNet`` (or ``from segmentation.segnet import SegNet`` /
``from generative.gan import GAN``).  ``load_reference_model`` makes those names resolve to the
shim namespace and to the facade's modules for the duration of the import, then restores
``sys.modules``.  The files themselves are read from a staged copy (``baseline/_ref``, created by
scripts/stage_reference.py, git-ignored) or from /root/reference when it exists; nothing of the
reference is vendored in the repository.
"""
import importlib.util
import os
import sys
import types

from . import tfshim

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEARCH_ROOTS = [os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]


def reference_root():
    for root in SEARCH_ROOTS:
        if os.path.isfile(os.path.join(root, "models", "resnet_v1_5.py")):
            return root
    return None


def _tf_package():
    """A fake ``tensorflow`` package whose ``compat.v1`` is the shim."""
    pkg = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    pkg.compat = compat
    compat.v1 = tfshim
    pkg.__path__ = []
    compat.__path__ = []
    return {"tensorflow": pkg, "tensorflow.compat": compat, "tensorflow.compat.v1": tfshim}


def load_reference_model(rel_path, facade_modules, root=None, quiet=True):
    """Import ``<root>/<rel_path>`` (e.g. 'models/resnet_v1_5.py') and return the module.

    facade_modules: {'convnet': module[, 'segmentation.segnet': module, 'generative.gan': module]}
    — the engine's implementations of the bases the model files import.
    """
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError(
            "reference model files not found; run scripts/stage_reference.py where /root/reference "
            "is available (searched: %s)" % ", ".join(SEARCH_ROOTS))
    path = os.path.join(root, rel_path)
    injected = dict(_tf_package())
    injected.update(facade_modules)
    for pkg in ("segmentation", "generative", "models"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        injected.setdefault(pkg, m)
    saved = {k: sys.modules.get(k) for k in injected}
    extra = []
    try:
        sys.modules.update(injected)
        # model files that import sibling model files (deeplabv3plus -> models.resnet_v1_5_dilated)
        src = open(path).read()
        for dep in ("resnet_v1_5_dilated", "resnet_v1_5"):
            key = "models." + dep
            if ("from models.%s import" % dep) in src and key not in sys.modules:
                sys.modules[key] = _exec_file(os.path.join(root, "models", dep + ".py"), key, quiet)
                extra.append(key)
        name = "mcn_ref_" + rel_path.replace("/", "_").replace(".py", "") + "_" + \
               str(id(facade_modules.get("convnet")))
        return _exec_file(path, name, quiet)
    finally:
        for k in extra:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _exec_file(path, name, quiet):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if quiet:
        # the model files print every layer shape while building; keep that off the bench output
        mod.print = lambda *a, **k: None
    return mod


def product_facade():
    from . import convnet, gan, segnet
    return {"convnet": convnet, "segmentation.segnet": segnet, "generative.gan": gan}
