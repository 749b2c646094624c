"""Stand-in for the two attributes of ``trimesh.load(path, process=False)`` the reference uses (stitch_lp.py:96-97: ``.vertices``,
``.faces``), installed by the launcher only when the real package is not importable."""
from types import SimpleNamespace


def load(path, process=False, **kwargs):
    from ..stitch import read_obj_geometry
    v, f = read_obj_geometry(path)
    return SimpleNamespace(vertices=v, faces=f)
