from .nn import *  # noqa: F401,F403
