from .sync_bn import *  # noqa: F401,F403
