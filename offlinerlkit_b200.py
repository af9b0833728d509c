"""Import shim: makes the directory ``offlinerl-kit_b200/`` importable as the
package ``offlinerlkit_b200`` (a hyphen cannot appear in a Python module name).

``import offlinerlkit_b200`` works whenever the repository root is on sys.path;
sub-modules resolve through ``__path__`` below (``offlinerlkit_b200.buffer`` ->
``offlinerl-kit_b200/buffer.py``).
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "offlinerl-kit_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
