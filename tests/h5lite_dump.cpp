// test helper: dump what the C++ reader of hybridmc_b200/csrc/host/h5lite.h sees in a file
// (tests/test_h5lite.py compiles it with g++ and compares with the Python parser)
#include "../hybridmc_b200/csrc/host/h5lite.h"
#include <cinttypes>

static void show(const char *kind, const std::string &name, const h5lite::Item &it) {
  uint64_t sum = 1469598103934665603ull; // FNV-1a over the raw bytes
  for (unsigned char c : it.bytes) sum = (sum ^ c) * 1099511628211ull;
  std::printf("%s %s %u [", kind, name.c_str(), unsigned(it.dtype));
  for (size_t k = 0; k < it.dims.size(); k++) std::printf(k ? ",%" PRIu64 : "%" PRIu64, it.dims[k]);
  std::printf("] %zu %016" PRIx64 "\n", it.bytes.size(), sum);
}

int main(int argc, char **argv) {
  if (argc != 2) return 2;
  h5lite::Reader r;
  if (!r.open(argv[1])) return 1;
  for (auto &kv : r.datasets) show("dataset", kv.first, kv.second);
  for (auto &kv : r.attrs) show("attr", kv.first, kv.second);
  return 0;
}
