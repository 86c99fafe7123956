"""Loads the reference's own PillarVFE / PointPillarScatter from the files staged in baseline/_ref/ (stage_reference.py).
Used only by bench.py's reference arms (CPU baseline and the eager-torch GPU comparator); never by the product path."""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("vfe_template.py", "pillar_vfe.py", "pointpillar_scatter.py"))


def load():
    """-> (PillarVFE, PointPillarScatter) classes of the reference, unmodified."""
    pkg = types.ModuleType("hgsf_refvfe")
    pkg.__path__ = []
    sys.modules["hgsf_refvfe"] = pkg

    def _load(name, fname):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("hgsf_refvfe.vfe_template", "vfe_template.py")       # pillar_vfe.py imports it relatively
    vfe = _load("hgsf_refvfe.pillar_vfe", "pillar_vfe.py")
    sc = _load("hgsf_refscatter", "pointpillar_scatter.py")
    return vfe.PillarVFE, sc.PointPillarScatter
