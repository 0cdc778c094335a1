"""Drop-in alias: `import nse` resolves to the sm_100a implementation in d4s.nse."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from d4s.nse import *  # noqa: E402,F401,F403
from d4s import nse as _impl  # noqa: E402

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})
