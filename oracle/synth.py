"""TEST INFRASTRUCTURE — alias of tcdiff_b200/synth.py (deterministic synthetic weights / inputs / noise banks).

The generators live next to the product because bench.py's measured arm needs them too and must not import anything
under oracle/.  They are loaded here BY PATH (not through the `tcdiff_b200` package), so the oracle and
oracle/make_golden.py keep working without the CUDA library and never import product code paths.
"""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tcdiff_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("oracle._synth_impl", _path)
_impl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_impl)
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
