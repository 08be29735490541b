from .google import *  # noqa: F401,F403
