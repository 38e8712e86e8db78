#include "backbone_ops.h"
