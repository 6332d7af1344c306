from .h5lite import *  # noqa: F401,F403
from .h5write import *  # noqa: F401,F403
from .nag_io import *  # noqa: F401,F403
