// epi_async.cuh -- mbarrier + TMA bulk-copy helpers (sm_100a) shared by the staged kernels.
#pragma once
#include "epi_device.cuh"

namespace epi {

EPI_DI void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
EPI_DI void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
EPI_DI void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion counted on `bar`
EPI_DI void bulk_load_row(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          (unsigned)__cvta_generic_to_shared(sdst)),
      "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

}  // namespace epi
