from overflow_b200.constants import *  # noqa: F401,F403
