// Stand-in for HighFive (HDF5 C++ wrapper, downloaded by the reference's CMake; not available offline), for oracle/_ref
// only: the result-file writer is outside the hot path.  Every call is accepted and writes nothing.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
namespace HighFive {
struct DataSpace {
  DataSpace() = default;
  std::vector<std::size_t> getDimensions() const { return {}; }
  template <class... A>
  DataSpace(A&&...) {}
  static DataSpace From(...) { return DataSpace(); }
};
struct Attribute {
  template <class T>
  void write(const T&) {}
  template <class T>
  void read(T&) const {}
};
struct DataSet {
  template <class T>
  void read(T&) const {}
  template <class T>
  void read(T*) const {}
  template <class T>
  void read_raw(T*) const {}
  DataSpace getSpace() const { return DataSpace(); }
  std::vector<std::size_t> getDimensions() const { return {}; }
  template <class T>
  void write_raw(const T*) {}
  template <class T>
  void write(const T&) {}
  template <class T>
  Attribute createAttribute(const std::string&, const T&) { return Attribute(); }
  template <class T>
  Attribute createAttribute(const std::string&, const DataSpace&) { return Attribute(); }
};
struct Group {
  Group createGroup(const std::string&) { return Group(); }
  Group getGroup(const std::string&) const { return Group(); }
  DataSet getDataSet(const std::string&) const { return DataSet(); }
  bool exist(const std::string&) const { return false; }
  template <class T>
  DataSet createDataSet(const std::string&, const DataSpace&) { return DataSet(); }
  template <class T>
  DataSet createDataSet(const std::string&, const T&) { return DataSet(); }
  template <class T>
  Attribute createAttribute(const std::string&, const T&) { return Attribute(); }
  template <class T>
  Attribute createAttribute(const std::string&, const DataSpace&) { return Attribute(); }
};
struct File : Group {
  enum : unsigned { ReadOnly = 0, ReadWrite = 1, Truncate = 2, Excl = 4, Create = 8, Overwrite = 2, OpenOrCreate = 9 };
  File() = default;
  File(const std::string&, unsigned = 0) {}
  void flush() {}
};
}  // namespace HighFive
