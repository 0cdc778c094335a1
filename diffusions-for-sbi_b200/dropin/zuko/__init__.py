"""Names of `zuko` that pickled reference models refer to, mapped onto d4s.nn (zuko itself is not required)."""
from . import nn  # noqa: F401
