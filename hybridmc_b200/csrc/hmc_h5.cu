// C-ABI around the dependency-free HDF5 writer (host/h5lite.h); host code only.
// Replaces the H5Cpp calls of src/writer.h, src/config.cc:18-60, src/main.cc:383-409 and
// src/snapshot.cc:14-68 for callers that hold the results in memory (the `hybridmc`
// executable links the header directly; the Python campaign driver comes through here).
#include "../../include/hybridmc_b200.h"
#include "hmc_internal.h"
#include "host/h5lite.h"

struct hmc_h5_file {
  h5lite::Writer w;
};

namespace {
template <class F> int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    hmc_set_global_error(e.what());
    return -1;
  }
}
std::vector<uint64_t> dims_of(int rank, const uint64_t *dims) {
  if (rank < 0 || rank > 32 || (rank > 0 && !dims)) throw std::runtime_error("h5: bad rank / dims");
  return std::vector<uint64_t>(dims, dims + rank);
}
void check(hmc_h5 f, const char *name, int dtype, const void *data, uint64_t nbytes) {
  if (!f || !name || !*name) throw std::runtime_error("h5: null file or empty name");
  if (dtype < HMC_H5_F64 || dtype > HMC_H5_VSTR) throw std::runtime_error("h5: unknown dtype");
  if (nbytes && !data) throw std::runtime_error("h5: null data");
}
} // namespace

extern "C" {

int hmc_h5_create(hmc_h5 *out) {
  return guarded([&] {
    if (!out) throw std::runtime_error("h5: null output pointer");
    *out = new hmc_h5_file;
  });
}

int hmc_h5_dataset(hmc_h5 f, const char *path, int dtype, int rank, const uint64_t *dims,
                   const void *data, uint64_t nbytes) {
  return guarded([&] {
    check(f, path, dtype, data, nbytes);
    f->w.dataset(path, uint8_t(dtype), dims_of(rank, dims), data ? data : "", size_t(nbytes));
  });
}

int hmc_h5_attribute(hmc_h5 f, const char *name, int dtype, int rank, const uint64_t *dims,
                     const void *data, uint64_t nbytes) {
  return guarded([&] {
    check(f, name, dtype, data, nbytes);
    f->w.attribute(name, uint8_t(dtype), dims_of(rank, dims), data ? data : "", size_t(nbytes));
  });
}

int hmc_h5_write(hmc_h5 f, const char *filename, int do_fsync) {
  return guarded([&] {
    if (!f || !filename) throw std::runtime_error("h5: null file or name");
    if (!f->w.write(filename, do_fsync != 0))
      throw std::runtime_error(std::string("h5: cannot write ") + filename);
  });
}

int hmc_h5_destroy(hmc_h5 f) {
  delete f;
  return 0;
}

} // extern "C"
