"""Drop-in alias: `import embedding_nets` resolves to the sm_100a implementation in d4s.embedding_nets."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from d4s.embedding_nets import *  # noqa: E402,F401,F403
from d4s import embedding_nets as _impl  # noqa: E402

globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})
