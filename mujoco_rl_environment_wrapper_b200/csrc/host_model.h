// host_model.h — host-side compiled model: named int32/float64 fields + name
// tables, and its serialisation into the packed blob of include/mjb_blob.h.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mjb_blob.h"

namespace mjb {

struct HostModel {
  // field registries (insertion order kept for a stable blob layout)
  std::vector<std::string> order;
  std::map<std::string, std::vector<int32_t>> i32;
  std::map<std::string, std::vector<double>> f64;
  // name tables per object type (MJB_OBJ_*); unnamed objects hold ""
  std::map<int, std::vector<std::string>> names;
  std::vector<uint8_t> blob;

  std::vector<int32_t>& I(const std::string& n) {
    if (!i32.count(n)) { order.push_back(n); }
    return i32[n];
  }
  std::vector<double>& F(const std::string& n) {
    if (!f64.count(n)) { order.push_back(n); }
    return f64[n];
  }
  void set_int(const std::string& n, int v) { I(n) = {v}; }
  void set_f64(const std::string& n, double v) { F(n) = {v}; }
  int get_int(const std::string& n) const {
    auto it = i32.find(n);
    if (it == i32.end() || it->second.empty()) throw std::runtime_error("model: missing int field " + n);
    return it->second[0];
  }
  const std::vector<int32_t>& Iv(const std::string& n) const {
    auto it = i32.find(n);
    if (it == i32.end()) throw std::runtime_error("model: missing field " + n);
    return it->second;
  }
  const std::vector<double>& Fv(const std::string& n) const {
    auto it = f64.find(n);
    if (it == f64.end()) throw std::runtime_error("model: missing field " + n);
    return it->second;
  }

  void pack() {
    const size_t nf = order.size();
    size_t off = sizeof(mjb_blob_header) + nf * sizeof(mjb_blob_field);
    off = (off + 7) & ~size_t(7);
    std::vector<mjb_blob_field> dir(nf);
    for (size_t k = 0; k < nf; k++) {
      const std::string& n = order[k];
      memset(&dir[k], 0, sizeof(mjb_blob_field));
      if (n.size() >= MJB_FIELD_NAME_LEN) throw std::runtime_error("field name too long: " + n);
      strncpy(dir[k].name, n.c_str(), MJB_FIELD_NAME_LEN - 1);
      size_t bytes;
      if (i32.count(n)) {
        dir[k].dtype = MJB_DTYPE_I32;
        dir[k].count = (int32_t)i32[n].size();
        bytes = i32[n].size() * 4;
      } else {
        dir[k].dtype = MJB_DTYPE_F64;
        dir[k].count = (int32_t)f64[n].size();
        bytes = f64[n].size() * 8;
      }
      dir[k].offset = (int64_t)off;
      off += (bytes + 7) & ~size_t(7);
    }
    blob.assign(off, 0);
    mjb_blob_header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, MJB_BLOB_MAGIC, 8);
    h.nfields = (int32_t)nf;
    h.total_bytes = (int64_t)off;
    memcpy(blob.data(), &h, sizeof(h));
    memcpy(blob.data() + sizeof(h), dir.data(), nf * sizeof(mjb_blob_field));
    for (size_t k = 0; k < nf; k++) {
      const std::string& n = order[k];
      if (i32.count(n)) {
        if (!i32[n].empty()) memcpy(blob.data() + dir[k].offset, i32[n].data(), i32[n].size() * 4);
      } else {
        if (!f64[n].empty()) memcpy(blob.data() + dir[k].offset, f64[n].data(), f64[n].size() * 8);
      }
    }
  }
};

// MJCF text -> compiled model (throws std::runtime_error with a message).
void compile_mjcf(const std::string& xml_text, HostModel& out);

}  // namespace mjb
