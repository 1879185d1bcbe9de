/* The drop-in boundary from plain C99: include/orb_b200.h + liborbb200.so, nothing else.
 * Prints one "name rc" line per entry point tried on a small synthetic image; on a machine without
 * a CUDA device every compute entry point must refuse with ORB_E_NODEVICE (there is no CPU path).
 *   gcc -std=c99 -Iinclude examples/abi_demo.c -Lorb_slam3_b200 -lorbb200 -Wl,-rpath,$PWD/orb_slam3_b200 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orb_b200.h"

int main(void) {
  enum { W = 320, H = 240, CAP = 700 };
  static uint8_t img[H * W];
  static orb_keypoint kps[CAP];
  static uint8_t desc[CAP * 32];
  unsigned s = 12345u;
  int i, n = 0, rc;
  orb_extractor* ex = NULL;
  orb_stereo* st = NULL;
  orb_poseopt* po = NULL;
  orb_frustum* fr = NULL;
  for (i = 0; i < H * W; i++) { s = s * 1664525u + 1013904223u; img[i] = (uint8_t)(s >> 24); }
  printf("version %s\n", orb_version());
  printf("devices %d\n", orb_device_count());
  rc = orb_create(500, 1.2f, 8, 20, 7, 0, &ex);
  printf("orb_create %d\n", rc);
  rc = orb_extract(ex, img, H, W, W, 0, 0, kps, desc, CAP, &n);
  printf("orb_extract %d n=%d err=\"%s\"\n", rc, n, rc < 0 ? orb_last_error() : "");
  rc = stereo_create(0, &st);
  printf("stereo_create %d\n", rc);
  rc = poseopt_create(0, &po);
  printf("poseopt_create %d\n", rc);
  rc = frustum_create(0, &fr);
  printf("frustum_create %d\n", rc);
  {
    uint8_t a[32], b[32];
    memset(a, 0x00, 32); memset(b, 0xff, 32);
    printf("ham_distance %d\n", ham_distance(a, b));
  }
  if (st) stereo_destroy(st);
  if (po) poseopt_destroy(po);
  if (fr) frustum_destroy(fr);
  orb_destroy(ex);
  return 0;
}
