from .graph import *  # noqa: F401,F403
