// TEST STUB — point types live in point_cloud.h of the stub tree.
#pragma once
#include "point_cloud.h"
