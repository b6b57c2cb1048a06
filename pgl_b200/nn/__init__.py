from . import functional  # noqa: F401
from .conv import *  # noqa: F401,F403
