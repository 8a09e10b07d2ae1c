"""Import shim: the package directory is named ``ffmpeg-ht_b200`` (hyphen), which Python cannot import by
name.  ``import ffmpeg_ht_b200`` loads that package from its directory."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "ffmpeg-ht_b200"
_spec = importlib.util.spec_from_file_location("ffmpeg_ht_b200", _pkg_dir / "__init__.py",
                                               submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ffmpeg_ht_b200"] = _mod
_spec.loader.exec_module(_mod)
