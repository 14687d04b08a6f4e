// XLA FFI custom-call shim over the C-ABI.
// jaxlib's FFI headers are not present in this image, so the handlers are
// compiled only where <xla/ffi/api/ffi.h> exists; elsewhere this TU is empty.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define SNNQP_HAVE_XLA_FFI 1
#endif
#endif

#ifdef SNNQP_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "snnqp.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  return rc == SNNQP_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInvalidArgument, snnqp_last_error());
}

// SpikingBlock(QuantConv3x3, BatchNorm, multi_step_LIF) [+ max-pool]; x: (B,T,H,W,Cin) uint8, batch-major
static ffi::Error SpikingConvImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::S8> wq,
                                  ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                  ffi::ResultBuffer<ffi::U8> spikes, int32_t pool, float tau, float v_th,
                                  float v_reset) {
  auto d = x.dimensions();
  snnqp_block_params p{};
  p.B = (int32_t)d[0]; p.T = (int32_t)d[1]; p.H = (int32_t)d[2]; p.W = (int32_t)d[3]; p.Cin = (int32_t)d[4];
  p.Cout = (int32_t)scale.dimensions()[0];
  p.x_stride_t = (int64_t)p.H * p.W * p.Cin; p.x_stride_b = p.x_stride_t * p.T;
  const int Ho = pool ? p.H / 2 : p.H, Wo = pool ? p.W / 2 : p.W;
  p.y_stride_t = (int64_t)Ho * Wo * p.Cout; p.y_stride_b = p.y_stride_t * p.T;
  p.tau = tau; p.v_threshold = v_th; p.v_reset = v_reset; p.pool = pool; p.impl = SNNQP_IMPL_AUTO;
  return Status(snnqp_spiking_conv3x3_fwd(&p, x.typed_data(), nullptr, wq.typed_data(), scale.typed_data(),
                                          bias.typed_data(), spikes->typed_data(), nullptr, nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(SnnqpSpikingConv, SpikingConvImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("pool")
                                  .Attr<float>("tau")
                                  .Attr<float>("v_threshold")
                                  .Attr<float>("v_reset"));

// SpikingBlock(QuantDense, multi_step_LIF); x: (B,T,K) uint8
static ffi::Error SpikingDenseImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::S8> wq,
                                   ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                                   ffi::ResultBuffer<ffi::U8> spikes, float tau, float v_th, float v_reset) {
  auto d = x.dimensions();
  snnqp_block_params p{};
  p.B = (int32_t)d[0]; p.T = (int32_t)d[1]; p.H = p.W = 1; p.Cin = (int32_t)d[2];
  p.Cout = (int32_t)scale.dimensions()[0];
  p.x_stride_t = p.Cin; p.x_stride_b = (int64_t)p.T * p.Cin;
  p.y_stride_t = p.Cout; p.y_stride_b = (int64_t)p.T * p.Cout;
  p.tau = tau; p.v_threshold = v_th; p.v_reset = v_reset; p.impl = SNNQP_IMPL_AUTO;
  return Status(snnqp_spiking_dense_fwd(&p, x.typed_data(), nullptr, wq.typed_data(), scale.typed_data(),
                                        bias.typed_data(), spikes->typed_data(), nullptr, nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(SnnqpSpikingDense, SpikingDenseImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<float>("tau")
                                  .Attr<float>("v_threshold")
                                  .Attr<float>("v_reset"));

// TCJA (models.py:41-95) from the spike counts a conv block accumulated; att: (B,T,C) fp32
static ffi::Error TcjaImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> counts, ffi::Buffer<ffi::S8> wq_t,
                           ffi::Buffer<ffi::S8> wq_c, ffi::Buffer<ffi::F32> scale_t, ffi::Buffer<ffi::F32> scale_c,
                           ffi::ResultBuffer<ffi::F32> att, int32_t spatial) {
  auto d = counts.dimensions();                 // (B, T, C)
  snnqp_block_params p{};
  p.B = (int32_t)d[0]; p.T = (int32_t)d[1]; p.Cin = p.Cout = (int32_t)d[2]; p.H = p.W = spatial;
  p.att_stride_t = p.Cin; p.att_stride_b = (int64_t)p.T * p.Cin; p.att_mod = p.Cin;
  return Status(snnqp_tcja_fwd(&p, nullptr, wq_t.typed_data(), wq_c.typed_data(), scale_t.typed_data(),
                               scale_c.typed_data(), counts.typed_data(), att->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(SnnqpTcja, TcjaImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("spatial"));

// vote (models.py:253-255): spikes (B,T,N) uint8 -> logits (B, N / group) fp32
static ffi::Error VoteImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> spikes, ffi::ResultBuffer<ffi::F32> logits,
                           int32_t group) {
  auto d = spikes.dimensions();
  return Status(snnqp_vote_fwd(spikes.typed_data(), (int)d[1], (int)d[0], (int)d[2], group, d[2], d[1] * d[2],
                               logits->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(SnnqpVote, VoteImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("group"));

// event -> frame integration (examples/input_pipeline.py:142-219): addrs (N,3) int32, offsets (B+1) int64
static ffi::Error EventsToFramesImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> addrs, ffi::Buffer<ffi::S64> offsets,
                                     ffi::ResultBuffer<ffi::U8> frames, int32_t sensor_wh, int32_t resolution_scale,
                                     int64_t max_events_per_sample) {
  auto f = frames->dimensions();                // (B, T, wh, wh, 2)
  return Status(snnqp_events_to_frames(addrs.typed_data(), offsets.typed_data(), (int)f[0], (int)f[1], sensor_wh,
                                       resolution_scale, max_events_per_sample, frames->typed_data(), 0, nullptr,
                                       stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(SnnqpEventsToFrames, EventsToFramesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("sensor_wh")
                                  .Attr<int32_t>("resolution_scale")
                                  .Attr<int64_t>("max_events_per_sample"));
#endif  // SNNQP_HAVE_XLA_FFI
