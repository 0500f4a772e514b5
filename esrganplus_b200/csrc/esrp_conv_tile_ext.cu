#define ESRP_EXT true
#define ESRP_PLAN_TILE_NAME plan_tile_ext
#include "plan_tile.inl"
