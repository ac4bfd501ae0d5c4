/* C restatement of the AudioMPS (PsiCMPS) scan of /root/reference/model.py -- TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under audio_mps_b200/ links or calls this; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py do, as the checker / the CPU thing that is timed.
 *
 * PARITY UNPINNED: the reference is TensorFlow-1.x Python and holds no golden numbers; this file is
 * pinned against oracle/cmps_oracle.py (the op-for-op PyTorch restatement) by
 * tests/test_oracle_golden.py and, through it, against the minted fixtures in tests/golden/.
 *
 * Built twice from cmps_ref_impl.h: float32 arithmetic (the reference's) and float64 arithmetic
 * with the reference's float32 definitions kept (t_k running sum, fl32(f*t_k) phase angle).
 * OpenMP over clips.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* t_0 = 0, t_{k+1} = fl32(t_k + fl32(delta_t))   (model.py:16,157,281) */
static void ttable(int n, double delta_t, float* tt) {
  volatile float t = 0.0f;
  const float dt = (float)delta_t;
  for (int k = 0; k < n; ++k) {
    tt[k] = t;
    t = t + dt;
  }
}

#define REAL float
#define SUF f32
#define IS_F32 1
#include "cmps_ref_impl.h"
#undef REAL
#undef SUF
#undef IS_F32

#define REAL double
#define SUF f64
#define IS_F32 0
#include "cmps_ref_impl.h"
#undef REAL
#undef SUF
#undef IS_F32

int cmps_ref_abi(void) { return 1; }
