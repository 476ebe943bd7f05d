/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Plain-C restatement of the GAE-lambda recursion of the rollout storage (upstream allenact
 * RolloutBlockStorage.compute_returns, which the SafeVLA fork runs once on (rewards, value_preds) and once on
 * (costs, c_value_preds); call-site witness architecture/models/allenact_transformer_models/inference_agent.py:255-267).
 * Same operation order as the sequential recursion, fp32 throughout, no FMA contraction (built with
 * -ffp-contract=off): the independent second opinion behind the "bit-exact" claim of the CUDA march kernel, which is
 * compared against oracle/torch_oracle.py::gae_returns -- and that restatement is compared against this file
 * (tests/test_cpu.py::test_gae_c_restatement_is_bit_identical).
 *
 *   delta_t = r_t + gamma * V_{t+1} * m_{t+1} - V_t
 *   g_t     = delta_t + (gamma * lambda) * m_{t+1} * g_{t+1}
 *   ret_t   = g_t + V_t ,   ret_T = V_T ,   adv_t = ret_t - V_t
 *
 * rewards [T, N]; value_preds, masks [T + 1, N]; returns [T + 1, N]; adv [T, N].  parity: unpinned upstream (the fork
 * is not vendored); pinned on closed forms and on the torch restatement. */
#include <stddef.h>

void gae_returns_f32(const float* rewards, const float* value_preds, const float* masks, float* returns, float* adv,
                     int T, int N, double gamma, double lam) {
  const float g32 = (float)gamma;          /* python scalar * fp32 tensor rounds the scalar to fp32 */
  const float gl32 = (float)(gamma * lam); /* gamma * lam is a python double product first */
  for (int n = 0; n < N; ++n) {
    float g = 0.0f;
    returns[(size_t)T * N + n] = value_preds[(size_t)T * N + n];
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * N + n, j = (size_t)(t + 1) * N + n;
      const float delta = rewards[i] + g32 * value_preds[j] * masks[j] - value_preds[i];
      g = delta + gl32 * masks[j] * g;
      returns[i] = g + value_preds[i];
      adv[i] = returns[i] - value_preds[i];
    }
  }
}

/* discounted returns without GAE (use_gae = False): ret_t = ret_{t+1} * gamma * m_{t+1} + r_t */
void discounted_returns_f32(const float* rewards, const float* value_preds, const float* masks, float* returns, int T,
                            int N, double gamma) {
  const float g32 = (float)gamma;
  for (int n = 0; n < N; ++n) {
    returns[(size_t)T * N + n] = value_preds[(size_t)T * N + n];
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * N + n, j = (size_t)(t + 1) * N + n;
      returns[i] = returns[j] * g32 * masks[j] + rewards[i];
    }
  }
}
