// Host-side check of csrc/ppb_cairo.cu's per-pixel scan conversion (tests/test_cairo_oracle.py builds and runs it
// where there is no GPU):  sprite_host_check r s_thick fill8hex stroke8hex  -> sz, then sz*sz pixels as hex words.
#define PPB_CAIRO_HOST_CHECK 1
#include "../pyphy_engine_b200/csrc/ppb_cairo.cu"
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const int r = atoi(argv[1]), st = atoi(argv[2]);
  const uint32_t f = (uint32_t)strtoul(argv[3], nullptr, 16), s = (uint32_t)strtoul(argv[4], nullptr, 16);
  std::vector<uint32_t> out((size_t)4 * (r + st) * (r + st) + 1);
  const int sz = ppb::host_check_sprite(r, st, f, s, out.data());
  if (sz < 0) return 3;
  printf("%d\n", sz);
  for (int i = 0; i < sz * sz; ++i) printf("%08x\n", out[i]);
  return 0;
}
