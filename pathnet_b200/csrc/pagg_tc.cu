// Tensor-core (tcgen05) math mode of the PAGG LSTM — under construction.
#include "pagg_common.cuh"

namespace pn {
int pagg_tc_forward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm,
                         const int32_t *seq_node, const uint8_t *seq_type, cudaStream_t st)
{
    (void)c; (void)w; (void)prm; (void)seq_node; (void)seq_type; (void)st;
    set_error("PN_MATH_TC: tensor-core LSTM not built in this library");
    return PN_ERR_UNSUPPORTED;
}
}  // namespace pn
