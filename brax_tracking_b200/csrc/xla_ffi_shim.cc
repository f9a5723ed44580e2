// XLA typed-FFI shim (SOURCE ONLY in this image: jaxlib / xla/ffi/api/ffi.h are not installed).
//
// This is the thin layer BASELINE.json asks for: a JAX host keeps `State` as device arrays and calls
//   jax.ffi.ffi_call("bt_step", ...)(action, qpos, qvel, act, qacc_warmstart, time, xpos, first_*, done, info_f, info_i)
// inside its jit / scan exactly where /root/reference/custom_brax/custom_wrappers.py:54-80 calls env.step; XLA hands
// the buffers (batch-leading [n_envs, dim] row-major, i.e. the layout vmap already uses) and its compute stream to the
// C ABI of include/bt_api.h.  Inputs that the kernel updates in place are declared as input/output aliases on the
// Python side (input_output_aliases), mirroring XLA buffer donation of the reference's functional State.
//
// Build (on a machine with jaxlib):
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//       brax_tracking_b200/csrc/xla_ffi_shim.cc -Lbrax_tracking_b200 -lbt_b200 -o libbt_xla_ffi.so
// Register:  jax.ffi.register_ffi_target("bt_step", jax.ffi.pycapsule(lib.BtStepFfi), platform="CUDA")
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include "bt_api.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error BtStepImpl(cudaStream_t stream, int64_t model_handle, ffi::Buffer<ffi::F32> action,
                             ffi::Buffer<ffi::F32> first_qpos, ffi::Buffer<ffi::F32> first_qvel, ffi::Buffer<ffi::F32> first_act,
                             ffi::Buffer<ffi::F32> first_warm, ffi::Buffer<ffi::F32> first_time, ffi::Buffer<ffi::F32> first_xpos,
                             ffi::Buffer<ffi::F32> first_obs, ffi::Buffer<ffi::S32> first_info_i,
                             // aliased in/out (donated) buffers
                             ffi::ResultBuffer<ffi::F32> qpos, ffi::ResultBuffer<ffi::F32> qvel, ffi::ResultBuffer<ffi::F32> act,
                             ffi::ResultBuffer<ffi::F32> warm, ffi::ResultBuffer<ffi::F32> time, ffi::ResultBuffer<ffi::F32> xpos,
                             ffi::ResultBuffer<ffi::F32> done, ffi::ResultBuffer<ffi::F32> info_f, ffi::ResultBuffer<ffi::S32> info_i,
                             // pure outputs
                             ffi::ResultBuffer<ffi::F32> obs, ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> metrics) {
  BtModel* m = reinterpret_cast<BtModel*>(model_handle);
  const int n = static_cast<int>(action.dimensions()[0]);
  BtStatePtrs st = {qpos->typed_data(), qvel->typed_data(), act->typed_data(), warm->typed_data(), time->typed_data(), xpos->typed_data()};
  BtStatePtrs first = {first_qpos.typed_data(), first_qvel.typed_data(), first_act.typed_data(), first_warm.typed_data(),
                       first_time.typed_data(), first_xpos.typed_data()};
  const int rc = bt_step(m, n, action.typed_data(), st, first, first_obs.typed_data(), first_info_i.typed_data(), obs->typed_data(),
                         reward->typed_data(), done->typed_data(), metrics->typed_data(), info_f->typed_data(), info_i->typed_data(),
                         stream);
  if (rc != BT_OK) return ffi::Error(ffi::ErrorCode::kInternal, bt_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BtStepFfi, BtStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // action
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());
#endif
