/* moved to include/h264gpu_mb_syntax.h (it is part of the C-ABI now) */
#include "h264gpu_mb_syntax.h"
