/* Minimal C consumer of libdm_b200.so (include/dm_b200.h): what a non-Python host links against.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Limage_compression_analysis_b200 -ldm_b200 \
 *       -Wl,-rpath,$PWD/image_compression_analysis_b200 -o /tmp/dm_example && /tmp/dm_example
 *
 * Without a GPU it reports the library's error text and exits 0 (there is no CPU path to fall back to);
 * with one it runs dm_fused_stats on a tiny synthetic pair whose device buffers the caller would normally own
 * through its own CUDA allocations -- here the example only checks the boundary, so it stops after the queries. */
#include <stdio.h>

#include "dm_b200.h"

int main(void) {
  printf("libdm_b200 ABI %d (header %d)\n", dm_abi_version(), DM_ABI_VERSION);
  if (dm_abi_version() != DM_ABI_VERSION) return 1;
  int sms = dm_device_sm_count();
  if (sms < 0) {
    printf("no usable CUDA device: %s\n", dm_last_error());
    return 0;
  }
  printf("%d SMs, %lld kernels launched so far, workspace %lld bytes, %d Sobel / %d SSIM partial slots per band\n", sms,
         (long long)dm_launch_count(), (long long)dm_workspace_bytes(), dm_sobel_nblocks(), dm_ssim_nblocks());
  /* argument checking is part of the contract: a null pair is DM_EARG with a message, not a crash */
  if (dm_fused_stats(NULL, NULL, DM_VALID_METRICS, 0, 0, NULL, NULL, NULL, NULL) != DM_EARG) return 2;
  printf("dm_fused_stats(NULL, ...) -> DM_EARG: %s\n", dm_last_error());
  return 0;
}
