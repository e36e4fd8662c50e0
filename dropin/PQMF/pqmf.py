from pqmf_b200._api import *  # noqa: F401,F403
from pqmf_b200._api import __all__  # noqa: F401
