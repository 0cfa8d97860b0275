import sys

FLOAT_MAX = sys.float_info.max
MINUS_FLOAT_MAX = -FLOAT_MAX
