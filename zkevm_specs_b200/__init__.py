"""Import alias: the package directory is `zkevm-specs_b200/` (not a valid Python
identifier), so `import zkevm_specs_b200` resolves to that directory through this stub."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "zkevm-specs_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
