from .graph import *  # noqa: F401,F403
from .sampling import *  # noqa: F401,F403
