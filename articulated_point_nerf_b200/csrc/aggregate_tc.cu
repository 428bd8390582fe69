// placeholder until the tcgen05 path lands (replaced in a later commit of this round)
#include "common.cuh"
extern "C" size_t apn_aggregate_tc_weights_bytes(int d_in) { (void)d_in; return 0; }
extern "C" int apn_aggregate_tc_pack_weights(const apn_mlp_weights* w, int d_in, void* packed, apn_stream_t stream) {
  (void)w; (void)d_in; (void)packed; (void)stream;
  apn_set_error("apn_aggregate_tc_pack_weights: tcgen05 path not built");
  return -3;
}
extern "C" int apn_aggregate_fwd_tc(const apn_agg_inputs* in, const apn_mlp_weights* w, const void* packed_weights,
                                    const apn_agg_outputs* out, int precision, apn_stream_t stream) {
  (void)in; (void)w; (void)packed_weights; (void)out; (void)precision; (void)stream;
  apn_set_error("apn_aggregate_fwd_tc: tcgen05 path not built");
  return -3;
}
