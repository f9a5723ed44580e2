// XLA typed-FFI shim (SOURCE ONLY in this image: jaxlib / xla/ffi/api/ffi.h are not installed; the CPU test-suite
// type-checks it against a stand-in of the binder, tests/test_abi.py::test_xla_ffi_shim_type_checks).
//
// This is the thin layer BASELINE.json asks for: a JAX host keeps `State` as device arrays and calls
//   jax.ffi.ffi_call("bt_reset", ...)(keys)                                    where custom_ppo.py:220-223 calls env.reset
//   jax.ffi.ffi_call("bt_step", ...)(action, first_*..., qpos, ..., info_i)    where custom_wrappers.py:54-80 calls env.step
// inside its jit / scan; XLA hands the buffers (batch-leading [n_envs, dim] row-major, i.e. the layout vmap already uses) and
// its compute stream to the C ABI of include/bt_api.h.  The model handle (BtModel*, from bt_model_create called once through
// ctypes when the env object is built) travels as an int64 attribute.
//
// In/out state: the nine buffers bt_step updates in place are passed as operands AND declared as results with
// input_output_aliases = {9: 0, 10: 1, ..., 17: 8} (operand index -> result index), which mirrors XLA's donation of the reference's
// functional State.  An aliased operand must still be bound as an Arg: XLA hands the handler both views.  When XLA could not
// alias (the operand is still live elsewhere) the two pointers differ and the handler copies operand -> result first.
//
// Build (on a machine with jaxlib):
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude -I/usr/local/cuda/include \
//       brax_tracking_b200/csrc/xla_ffi_shim.cc -Lbrax_tracking_b200 -lbt_b200 -lcudart -o libbt_xla_ffi.so
// Register:  jax.ffi.register_ffi_target("bt_step", jax.ffi.pycapsule(lib.BtStepFfi), platform="CUDA")   (same for bt_reset)
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include "bt_api.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
typedef ffi::Buffer<ffi::F32> F32Buf;
typedef ffi::Buffer<ffi::S32> S32Buf;
typedef ffi::ResultBuffer<ffi::F32> F32Res;
typedef ffi::ResultBuffer<ffi::S32> S32Res;

// operand -> aliased result: nothing to do when XLA aliased them, a device copy otherwise
template <typename In, typename Out>
static bool alias_or_copy(cudaStream_t stream, const In& in, Out& out) {
  if ((const void*)in.typed_data() == (const void*)out->typed_data()) return true;
  return cudaMemcpyAsync(out->typed_data(), in.typed_data(), in.size_bytes(), cudaMemcpyDeviceToDevice, stream) == cudaSuccess;
}

// wrap(env).step -- custom_brax/custom_wrappers.py:54-80 o EpisodeWrapper.step o envs/fruitfly.py:497-596
static ffi::Error BtStepImpl(cudaStream_t stream, int64_t model_handle, F32Buf action,
                             // the cached first state of the auto-reset wrapper (custom_wrappers.py:46-52): read only
                             F32Buf first_qpos, F32Buf first_qvel, F32Buf first_act, F32Buf first_warm, F32Buf first_time,
                             F32Buf first_xpos, F32Buf first_obs, S32Buf first_info_i,
                             // operands 9..17: the state the step advances (aliased to results 0..8)
                             F32Buf qpos_in, F32Buf qvel_in, F32Buf act_in, F32Buf warm_in, F32Buf time_in, F32Buf xpos_in,
                             F32Buf done_in, F32Buf info_f_in, S32Buf info_i_in,
                             // operand 18: the clip of every environment (RodentMultiClip; zeros for a single-clip model), read only
                             S32Buf clip_idx,
                             // results 0..8 (aliased) and 9..11 (pure outputs)
                             F32Res qpos, F32Res qvel, F32Res act, F32Res warm, F32Res time, F32Res xpos, F32Res done, F32Res info_f,
                             S32Res info_i, F32Res obs, F32Res reward, F32Res metrics) {
  BtModel* m = reinterpret_cast<BtModel*>(model_handle);
  const int n = static_cast<int>(action.dimensions()[0]);
  bool ok = alias_or_copy(stream, qpos_in, qpos) && alias_or_copy(stream, qvel_in, qvel) && alias_or_copy(stream, act_in, act) &&
            alias_or_copy(stream, warm_in, warm) && alias_or_copy(stream, time_in, time) && alias_or_copy(stream, xpos_in, xpos) &&
            alias_or_copy(stream, done_in, done) && alias_or_copy(stream, info_f_in, info_f) && alias_or_copy(stream, info_i_in, info_i);
  if (!ok) return ffi::Error(ffi::ErrorCode::kInternal, "bt_step: operand -> result copy failed");
  BtStatePtrs st = {qpos->typed_data(), qvel->typed_data(), act->typed_data(), warm->typed_data(), time->typed_data(), xpos->typed_data()};
  BtStatePtrs first = {first_qpos.typed_data(), first_qvel.typed_data(), first_act.typed_data(), first_warm.typed_data(),
                       first_time.typed_data(), first_xpos.typed_data()};
  const int rc = bt_step(m, n, action.typed_data(), st, first, first_obs.typed_data(), first_info_i.typed_data(), obs->typed_data(),
                         reward->typed_data(), done->typed_data(), metrics->typed_data(), info_f->typed_data(), info_i->typed_data(),
                         clip_idx.typed_data(), stream);
  if (rc != BT_OK) return ffi::Error(ffi::ErrorCode::kInternal, bt_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BtStepFfi, BtStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Arg<F32Buf>()                                                              // 0 action
                                  .Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()  // 1..6 first state
                                  .Arg<F32Buf>().Arg<S32Buf>()                                                // 7 first_obs, 8 first_info_i
                                  .Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>().Arg<F32Buf>()  // 9..14 state (aliased)
                                  .Arg<F32Buf>().Arg<F32Buf>().Arg<S32Buf>()                                  // 15 done, 16 info_f, 17 info_i (aliased)
                                  .Arg<S32Buf>()                                                              // 18 clip_idx
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>()  // 0..5 state
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<S32Buf>()                                  // 6 done, 7 info_f, 8 info_i
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>());                                // 9 obs, 10 reward, 11 metrics

// wrap(env).reset -- envs/fruitfly.py:449-495 (+ envs/rodent.py:154-159) under custom_brax/custom_wrappers.py:46-52; with
// fixed_start_frame >= 0 RenderRolloutWrapperTracking.reset (custom_wrappers.py:85-125).  The caller keeps copies of the returned
// state / obs / info_i as the wrapper's first_* (a plain jnp copy on the JAX side).
static ffi::Error BtResetImpl(cudaStream_t stream, int64_t model_handle, int64_t fixed_start_frame, ffi::Buffer<ffi::U32> keys,
                              F32Res qpos, F32Res qvel, F32Res act, F32Res warm, F32Res time, F32Res xpos, F32Res obs, F32Res reward,
                              F32Res done, F32Res metrics, F32Res info_f, S32Res info_i, S32Res clip_idx) {
  BtModel* m = reinterpret_cast<BtModel*>(model_handle);
  const int n = static_cast<int>(keys.dimensions()[0]);
  BtStatePtrs st = {qpos->typed_data(), qvel->typed_data(), act->typed_data(), warm->typed_data(), time->typed_data(), xpos->typed_data()};
  const int rc = bt_reset(m, n, keys.typed_data(), static_cast<int>(fixed_start_frame), st, obs->typed_data(), reward->typed_data(),
                          done->typed_data(), metrics->typed_data(), info_f->typed_data(), info_i->typed_data(), clip_idx->typed_data(),
                          stream);
  if (rc != BT_OK) return ffi::Error(ffi::ErrorCode::kInternal, bt_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BtResetFfi, BtResetImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Attr<int64_t>("fixed_start_frame")
                                  .Arg<ffi::Buffer<ffi::U32>>()                                               // keys [n, 2]
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>()  // state
                                  .Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<F32Buf>().Ret<S32Buf>()  // obs reward done metrics info_f info_i
                                  .Ret<S32Buf>());                                                            // clip_idx (the clip drawn per environment)
#endif
