// TEST INFRASTRUCTURE (oracle/): stand-in for the HDF5 C++ API, which is absent
// from this image.  It exists only so that the UNMODIFIED reference sources under
// /root/reference/src compile into oracle/_ref/ (config.h:11 includes <H5Cpp.h>).
// Every call is a no-op: the oracle never needs the reference's file output,
// only its arithmetic.  Not part of the product; never linked into it.
#ifndef HMC_ORACLE_SHIM_H5CPP_H
#define HMC_ORACLE_SHIM_H5CPP_H

#include <cstddef>
#include <cstdint>
#include <string>

typedef unsigned long long hsize_t;
static const hsize_t H5S_UNLIMITED = ~hsize_t(0);
enum { H5S_SELECT_SET = 0 };
enum { H5F_ACC_TRUNC = 2, H5F_ACC_RDONLY = 0 };
enum { H5F_SCOPE_GLOBAL = 1 };
enum { H5S_SCALAR_TAG = 0 };
static const size_t H5T_VARIABLE = ~size_t(0);

namespace H5 {

struct DataType {};
struct PredType : DataType {
  static inline const DataType NATIVE_DOUBLE{}, NATIVE_FLOAT{}, NATIVE_UINT{},
      NATIVE_UINT64{}, NATIVE_INT{};
};
struct StrType : DataType {
  StrType() {}
  StrType(int, size_t) {}
};

struct DataSpace {
  DataSpace() {}
  DataSpace(int, const hsize_t *, const hsize_t * = nullptr) {}
  void setExtentSimple(int, const hsize_t *, const hsize_t * = nullptr) {}
  void selectHyperslab(int, const hsize_t *, const hsize_t *) {}
  int getSimpleExtentDims(hsize_t *d) const {
    d[0] = 0;
    return 0;
  }
};
static const DataSpace H5S_SCALAR_SPACE;
#define H5S_SCALAR H5::H5S_SCALAR_SPACE

struct DSetCreatPropList {
  void setChunk(int, const hsize_t *) {}
  void setDeflate(int) {}
};

struct Attribute {
  void write(const DataType &, const void *) {}
  void read(const DataType &, void *) {}
};

struct DataSet {
  void write(const void *, const DataType &, const DataSpace & = DataSpace(),
             const DataSpace & = DataSpace()) {}
  void write(const std::string &, const DataType &) {}
  void read(void *, const DataType &, const DataSpace & = DataSpace(),
            const DataSpace & = DataSpace()) const {}
  void read(std::string &, const DataType &, const DataSpace & = DataSpace()) const {}
  void extend(const hsize_t *) {}
  DataSpace getSpace() const { return DataSpace(); }
};

struct H5Object {
  Attribute createAttribute(const std::string &, const DataType &,
                            const DataSpace &) {
    return Attribute();
  }
  Attribute openAttribute(const std::string &) const { return Attribute(); }
};

struct Group;
struct H5Location : H5Object {
  DataSet createDataSet(const std::string &, const DataType &, const DataSpace &,
                        const DSetCreatPropList & = DSetCreatPropList()) {
    return DataSet();
  }
  DataSet openDataSet(const std::string &) const { return DataSet(); }
  inline Group createGroup(const std::string &);
  inline Group openGroup(const std::string &) const;
};
struct Group : H5Location {};
inline Group H5Location::createGroup(const std::string &) { return Group(); }
inline Group H5Location::openGroup(const std::string &) const { return Group(); }

struct H5File : H5Location {
  H5File() {}
  H5File(const std::string &, unsigned) {}
  void flush(int) {}
  void getVFDHandle(void **) {}
  void close() {}
};

} // namespace H5
#endif
