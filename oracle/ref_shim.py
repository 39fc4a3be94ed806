"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* Python-2 reference.

This file is part of ``oracle/``: it may be imported by ``tests/``, by
``oracle/make_golden.py`` and by nothing on the product path.

It runs the reference physics (``/root/reference/{geometry,dynamics,primitives,
dataio}.py``) in this container's Python 3 by
  * registering stub modules for the libraries the reference imports but that
    are absent here (``cairo``, ``glog``, ``matplotlib``, ``scipy.misc``),
  * an import hook that reads each reference module from disk and rewrites the
    Python-2-only syntax *in memory* (``print x`` statements, ``.iteritems()``,
    ``raw_input``); no reference source is copied into this repository,
  * patching ``World.generate_image`` to a no-op (pycairo is absent) and
    ``pdb.set_trace`` to raise, and wrapping ``Dynamics.resolve_collision`` so
    that every call is appended to an event log.

The reference directory does not exist on the GPU box, so nothing that runs
there imports this module; its outputs travel as ``tests/golden/*.npz``.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import re
import sys
import types

REF_DIR = os.environ.get("PYPHY_REFERENCE_DIR", "/root/reference")
_REF_MODULES = ("geometry", "physics", "primitives", "dynamics", "dataio",
                "save_data", "unit_tests")


class ReferenceTrap(Exception):
    """Raised where the reference would have dropped into ``pdb.set_trace()``."""


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "primitives.py"))


class _Swallow:
    """Object that accepts any attribute access / call (stands in for pycairo)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Swallow()

    def __call__(self, *a, **k):
        return _Swallow()


def _install_stubs() -> None:
    cairo = types.ModuleType("cairo")
    cairo.FORMAT_ARGB32 = 0
    cairo.Context = _Swallow

    class _ImageSurface:
        @staticmethod
        def create_for_data(*a, **k):
            return _Swallow()

    cairo.ImageSurface = _ImageSurface
    sys.modules.setdefault("cairo", cairo)
    sys.modules.setdefault("glog", types.ModuleType("glog"))
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        anim = types.ModuleType("matplotlib.animation")
        mpl.pyplot = plt
        mpl.animation = anim
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.animation"] = anim
    import scipy  # noqa: F401  (real scipy; only scipy.misc is missing)
    if "scipy.misc" not in sys.modules:
        misc = types.ModuleType("scipy.misc")
        misc.imsave = lambda *a, **k: None
        sys.modules["scipy.misc"] = misc
        import scipy as _sp
        _sp.misc = misc


_PRINT_RE = re.compile(r"^(\s*)print[ \t]+(?!\()(.*?)\s*$")
_PRINT_PAREN_RE = re.compile(r"^(\s*)print[ \t]*\((.*)\)\s*$")


def _py2_to_py3(src: str) -> str:
    out = []
    in_triple = False
    for line in src.split("\n"):
        # leave the bodies of ''' ... ''' blocks alone (they hold dead py2 code)
        if line.count("'''") % 2 == 1:
            in_triple = not in_triple
            out.append(line)
            continue
        if not in_triple:
            m = _PRINT_RE.match(line)
            if m and not _PRINT_PAREN_RE.match(line):
                line = "%sprint(%s)" % (m.group(1), m.group(2))
        out.append(line)
    src = "\n".join(out)
    src = src.replace(".iteritems()", ".items()")
    src = src.replace("raw_input(", "input(")
    return src


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname in _REF_MODULES:
            fn = os.path.join(REF_DIR, fullname + ".py")
            if os.path.isfile(fn):
                return importlib.util.spec_from_loader(fullname, self, origin=fn)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        fn = os.path.join(REF_DIR, module.__name__ + ".py")
        with open(fn, "r") as fh:
            src = _py2_to_py3(fh.read())
        module.__file__ = fn
        exec(compile(src, fn, "exec"), module.__dict__)


_loaded = None


def load():
    """Import the reference and return a namespace with its modules + event log.

    Returns an object with attributes ``gm``, ``dy``, ``pm``, ``dio`` (the
    reference's geometry/dynamics/primitives/dataio modules) and ``events`` (a
    list that every ``Dynamics.resolve_collision`` call appends
    ``(model, name, partner_name)`` to).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF_DIR)
    _install_stubs()
    # the product package ships façade modules of the same names; make sure the
    # hook wins for the bare names
    for name in _REF_MODULES:
        sys.modules.pop(name, None)
    finder = _RefFinder()
    sys.meta_path.insert(0, finder)
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            import geometry as gm
            import dynamics as dy
            import primitives as pm
            import dataio as dio
    finally:
        sys.meta_path.remove(finder)

    def _trap(*a, **k):
        raise ReferenceTrap("pdb.set_trace() reached in the reference")

    for mod in (gm, dy, pm, dio):
        if hasattr(mod, "pdb"):
            fake_pdb = types.SimpleNamespace(set_trace=_trap)
            mod.pdb = fake_pdb

    pm.World.generate_image = lambda self, *a, **k: None

    ns = types.SimpleNamespace(gm=gm, dy=dy, pm=pm, dio=dio, events=[])
    _orig_resolve = pm.Dynamics.resolve_collision

    def _logged_resolve(self, obj, name, deltaT):
        partner = self.objCol_[name][1]
        ns.events.append((self, name, partner))
        return _orig_resolve(self, obj, name, deltaT)

    pm.Dynamics.resolve_collision = _logged_resolve
    _loaded = ns
    return ns


def unload() -> None:
    """Drop the reference modules from ``sys.modules`` (so the product façade
    modules of the same names can be imported afterwards)."""
    global _loaded
    for name in _REF_MODULES:
        sys.modules.pop(name, None)
    _loaded = None
