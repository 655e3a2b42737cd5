"""`src` -- the drop-in surface of the reference's hot-path modules (see src/README.md).

The reference's `src` is a namespace package (it has no `__init__.py`) that also holds everything OUTSIDE the hot path
(`src.dataloader.*`, `src.custom_megapose.*`, `src.utils.bbox`, ...), which `test.py` needs and which this repository
does not re-implement.  A regular package would shadow all of that, so this package (and each of its sub-packages)
extends its `__path__` with the same-named directory of a reference checkout found on `sys.path` / `PYTHONPATH` (or
named by `GIGAPOSE_REFERENCE_ROOT`): modules that exist here win, everything else resolves from the checkout.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_src_dirs():
    """`<root>/src` directories of reference checkouts visible on sys.path (this repository's own excluded)."""
    found = []
    env = os.environ.get("GIGAPOSE_REFERENCE_ROOT")
    for entry in ([env] if env else []) + list(sys.path):
        if not entry:
            continue
        d = os.path.realpath(os.path.join(entry, "src"))
        if d != os.path.realpath(_HERE) and d not in found and os.path.isdir(os.path.join(d, "dataloader")):
            found.append(d)
    return found


def extend_path(path, name):
    """Appends the reference's directory for package `name` ('src', 'src.utils', ...) to `path` (a package __path__)."""
    rel = name.split(".")[1:]
    path = list(path)
    for root in reference_src_dirs():
        d = os.path.join(root, *rel)
        if os.path.isdir(d) and d not in path:
            path.append(d)
    return path


def reference_module(fullname):
    """The reference checkout's own copy of a module that also exists here (e.g. `src.utils.inout`, whose CNOS-detection
    loaders the dataloaders use), loaded under the alias `<fullname>__reference`; None without a checkout."""
    alias = fullname + "__reference"
    if alias in sys.modules:
        return sys.modules[alias]
    rel = fullname.split(".")[1:]
    for root in reference_src_dirs():
        f = os.path.join(root, *rel) + ".py"
        if os.path.isfile(f):
            spec = importlib.util.spec_from_file_location(alias, f)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                del sys.modules[alias]
                raise
            return mod
    return None


def fallback_getattr(module_name):
    """Module-level `__getattr__` (PEP 562) for a file that exists on both sides but is thinner here: names this
    repository does not provide are served from the reference checkout's copy of the same file."""
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        ref = reference_module(module_name)
        if ref is not None and hasattr(ref, name):
            return getattr(ref, name)
        raise AttributeError(f"module {module_name!r} has no attribute {name!r}")
    return __getattr__


__path__ = extend_path(__path__, __name__)
