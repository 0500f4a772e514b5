#define ESRP_EXT true
#define ESRP_PLAN_ROW_NAME plan_row_ext
#include "plan_row.inl"
