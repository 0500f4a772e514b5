#define ESRP_EXT false
#define ESRP_PLAN_ROW_NAME plan_row_base
#include "plan_row.inl"
