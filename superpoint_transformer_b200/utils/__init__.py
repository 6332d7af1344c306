from .nn import *  # noqa: F401,F403
from .tensor import *  # noqa: F401,F403
