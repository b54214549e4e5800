"""TEST INFRASTRUCTURE ONLY -- in-memory import of the *real* reference.

The reference (``/root/reference/arboris``) is Python 2.6 source.  It cannot be
imported by Python 3.12 / numpy 2.x as-is, and it must not be copied into this
repository.  This module installs a ``sys.meta_path`` finder that loads the
reference modules **from where they lie**, applying the seven mechanical
py3/numpy-2 fixes of SURVEY.md section 8(c) to the source text *in memory*
(nothing is written to disk, nothing of the reference enters the repo).

It only works where ``/root/reference`` exists (the build container).  On the
GPU box the reference is absent: tests then rely on the committed fixtures in
``tests/golden/`` that ``oracle/make_goldens.py`` generated with this loader.

Patches (reference file:line -> replacement), all arithmetic-neutral:

1. ``arboris/homogeneousmatrix.py:310``  ``p = H[0:3,3:4]`` (inside ``adjoint``)
   -> ``p = H[0:3,3]``  (numpy>=1.24 refuses the ragged ``array([[0,-p[2],..``)
2. ``arboris/core.py:100``   drop ``or isinstance(index, unicode)``
3. ``arboris/core.py:1081-1082``  ``itertools.imap`` -> builtin ``map``
4. ``arboris/core.py:260``   ``range(..)`` -> ``list(range(..))``
5. ``arboris/controllers.py:37-38``  wrap the lazy ``filter`` in ``list`` (on
   py3 the iterator is exhausted after one step and gravity vanishes)
6. ``arboris/robots/human36.py:106``  ``unicode(name)`` -> ``str(name)``
7. ``arboris/observers.py:241,262,269``  tab -> spaces, ``iterkeys/iteritems``
8. ``arboris/constraints.py:826``  the admissible eigenvalues (already filtered on
   ``S.imag == 0``) are taken as ``.real``: when *any* eigenvalue of B is complex
   ``eigvals`` returns a complex array and numpy 2 refuses the in-place
   ``A[0:3,0:3] -= s*diag(..)`` at :833 (2010-era numpy silently dropped the zero
   imaginary part).  SURVEY.md lists this as latent; the human36 contact scenario
   does hit it.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("ARBORIS_REFERENCE_ROOT", "/root/reference")


def _patch_homogeneousmatrix(src):
    # only the occurrence inside ``adjoint`` (the second one); ``inv`` needs (3,1)
    marker = "def adjoint(H):"
    head, tail = src.split(marker, 1)
    assert "p = H[0:3,3:4]" in tail
    tail = tail.replace("p = H[0:3,3:4]", "p = H[0:3,3]", 1)
    return head + marker + tail


def _patch_core(src):
    old = "if isinstance(index, str) or isinstance(index, unicode):"
    assert old in src
    src = src.replace(old, "if isinstance(index, str):")
    assert "from itertools import imap" in src
    src = src.replace("from itertools import imap", "imap = map")
    old = "self._dof = range(self._dof.start, self._dof.stop)"
    assert old in src
    src = src.replace(old, "self._dof = list(range(self._dof.start, self._dof.stop))")
    return src


def _patch_controllers(src):
    old = ("self._bodies = filter(lambda x: norm(x.mass>0.),\n"
           "                world.ground.iter_descendant_bodies())")
    assert old in src
    return src.replace(old, "self._bodies = list(filter(lambda x: norm(x.mass>0.),\n"
                       "                world.ground.iter_descendant_bodies()))")


def _patch_constraints(src):
    old = "S = S[logical_and(S.imag == 0, S.real <= 0)]"
    assert old in src
    return src.replace(old, old + ".real")


def _patch_human36(src):
    assert "name = unicode(name)" in src
    return src.replace("name = unicode(name)", "name = str(name)")


def _patch_observers(src):
    src = src.replace("\t", "        ")
    src = src.replace(".iterkeys()", ".keys()").replace(".iteritems()", ".items()")
    return src


_PATCHES = {
    "arboris.homogeneousmatrix": _patch_homogeneousmatrix,
    "arboris.core": _patch_core,
    "arboris.controllers": _patch_controllers,
    "arboris.constraints": _patch_constraints,
    "arboris.robots.human36": _patch_human36,
    "arboris.observers": _patch_observers,
}


class _PatchedLoader(importlib.machinery.SourceFileLoader):
    def get_data(self, path):  # source text is patched before compilation
        data = super().get_data(path)
        patch = _PATCHES.get(self.name)
        if patch is not None and path.endswith(".py"):
            data = patch(data.decode("utf-8")).encode("utf-8")
        return data

    def get_code(self, fullname):  # never read/write .pyc (tree is read-only)
        source = self.get_data(self.get_filename(fullname))
        return compile(source, self.get_filename(fullname), "exec", dont_inherit=True)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "arboris" and not fullname.startswith("arboris."):
            return None
        rel = fullname.split(".")
        base = os.path.join(REFERENCE_ROOT, *rel)
        if os.path.isdir(base):
            fn = os.path.join(base, "__init__.py")
            return importlib.util.spec_from_file_location(
                fullname, fn, loader=_PatchedLoader(fullname, fn),
                submodule_search_locations=[base])
        fn = base + ".py"
        if os.path.isfile(fn):
            return importlib.util.spec_from_file_location(
                fullname, fn, loader=_PatchedLoader(fullname, fn))
        return None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "arboris", "core.py"))


def install():
    """Make ``import arboris`` resolve to the patched-in-memory reference."""
    if not available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    import arboris  # noqa: F401
    return arboris
