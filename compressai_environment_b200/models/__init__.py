from .google import *  # noqa: F401,F403
from .waseda import *  # noqa: F401,F403
