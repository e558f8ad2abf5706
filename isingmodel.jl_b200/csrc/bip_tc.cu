// bip_tc.cu — placeholder until the tcgen05 path lands (next commit).
#include "handles.hpp"
namespace isb {
int bip_tc_model_init(isb_model *m, const double *) { return fail(m->ctx, ISB_ERR_UNSUPPORTED, "tensor-core path not built"); }
void bip_tc_model_free(isb_model *) {}
int bip_tc_ens_init(isb_ens *e) { return fail(e->model->ctx, ISB_ERR_UNSUPPORTED, "tensor-core path not built"); }
void bip_tc_ens_free(isb_ens *) {}
int bip_run_tc_device(isb_ens *e, int, int64_t, int, const double *, const double *, uint64_t, uint64_t, const double *,
                      int64_t, int64_t, double *) {
    return fail(e->model->ctx, ISB_ERR_UNSUPPORTED, "tensor-core path not built");
}
}  // namespace isb
