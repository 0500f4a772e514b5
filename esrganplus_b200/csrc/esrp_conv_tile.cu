#define ESRP_EXT false
#define ESRP_PLAN_TILE_NAME plan_tile_base
#include "plan_tile.inl"
