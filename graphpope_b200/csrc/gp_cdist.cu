// gp_cdist.cu — node2vec-branch pairwise block (reference utils.py:174-176).  Placeholder until
// the tcgen05 GEMM lands: reports GP_ERR_UNSUPPORTED so callers fail loudly.
#include "gp_internal.h"

extern "C" int gp_cdist_minmax(const float *, const float *, int64_t, int64_t, int64_t, int32_t, int32_t,
                               float *, int64_t, int64_t, gp_stream_t)
{
    gp_set_error("gp_cdist_minmax: not implemented in this build");
    return GP_ERR_UNSUPPORTED;
}
