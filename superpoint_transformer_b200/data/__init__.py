from .cluster import *  # noqa: F401,F403
from .data import *  # noqa: F401,F403
from .nag import *  # noqa: F401,F403
