"""Drop-in modules: the reference's operator API (module names, class/function signatures, state-dict
layouts) re-implemented on librnr_b200.so.  ``install()`` registers them in ``sys.modules`` under the
reference's top-level names so that train_rnr.py / test_rnr.py / train_dnr.py / test_dnr.py import
them unchanged (see relightable_nr_b200/run.py and INTEGRATION.md)."""
import importlib
import sys

_NAMES = {
    'misc': '.misc',
    'camera': '.camera',
    'render': '.render',
    'sph_harm': '.sph_harm',
    'network': '.network',
    'pytorch_prototyping': '.pytorch_prototyping_pkg',
    'pytorch_prototyping.pytorch_prototyping': '.pytorch_prototyping',
    'neural_renderer': '.neural_renderer',
    'gcn_lib': '.gcn_lib',
    'gcn_lib.dense': '.gcn_lib.dense',
    'gcn_lib.sparse': '.gcn_lib.sparse',
}


def install(names=None):
    """Register the drop-in modules under the reference's import names.  Returns the list installed."""
    done = []
    for top, rel in _NAMES.items():
        if names is not None and top not in names:
            continue
        try:
            mod = importlib.import_module(rel, __name__)
        except ModuleNotFoundError as e:      # a sub-package not built yet
            if e.name and e.name.startswith(__name__):
                continue
            raise
        sys.modules[top] = mod
        done.append(top)
    return done
