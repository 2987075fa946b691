// Memory-mapped safetensors reader (C ABI).  Format: u64-LE header length | JSON header | raw little-endian tensor data.
// The reference gets this from the Rust `safetensors` crate through transformers (SURVEY.md §2.2 native census).
// The JSON subset used by safetensors headers (objects, strings, integer arrays) is parsed in place.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#define ODB_API extern "C" __attribute__((visibility("default")))

namespace {

struct TensorInfo {
  std::string name, dtype;
  std::vector<int64_t> shape;
  uint64_t begin = 0, end = 0;
};

struct File {
  int fd = -1;
  uint8_t* map = nullptr;
  size_t size = 0;
  uint64_t data_off = 0;
  std::vector<TensorInfo> tensors;
  std::string error;
};

struct Parser {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  std::string str() {
    ws();
    std::string s;
    if (p >= e || *p != '"') { ok = false; return s; }
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) { ++p; }
      s.push_back(*p++);
    }
    if (p < e) ++p; else ok = false;
    return s;
  }
  int64_t num() {
    ws();
    bool neg = false;
    if (p < e && *p == '-') { neg = true; ++p; }
    int64_t v = 0;
    if (p >= e || *p < '0' || *p > '9') { ok = false; return 0; }
    while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (*p++ - '0');
    return neg ? -v : v;
  }
  void skip_value() {  // strings / numbers / nested objects / arrays (for __metadata__)
    ws();
    if (p >= e) { ok = false; return; }
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = (*p == '{') ? '}' : ']';
      int depth = 0;
      while (p < e) {
        if (*p == '"') { str(); continue; }
        if (*p == open) ++depth;
        else if (*p == close) { if (--depth == 0) { ++p; return; } }
        ++p;
      }
      ok = false;
      return;
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
  }
};

bool parse_header(File& f, const char* js, size_t n) {
  Parser ps{js, js + n};
  if (!ps.eat('{')) return false;
  while (ps.ok) {
    ps.ws();
    if (ps.eat('}')) break;
    std::string key = ps.str();
    if (!ps.eat(':')) return false;
    if (key == "__metadata__") {
      ps.skip_value();
    } else {
      TensorInfo t;
      t.name = key;
      if (!ps.eat('{')) return false;
      while (ps.ok) {
        if (ps.eat('}')) break;
        std::string k = ps.str();
        if (!ps.eat(':')) return false;
        if (k == "dtype") t.dtype = ps.str();
        else if (k == "shape") {
          if (!ps.eat('[')) return false;
          while (ps.ok && !ps.eat(']')) { t.shape.push_back(ps.num()); ps.eat(','); }
        } else if (k == "data_offsets") {
          if (!ps.eat('[')) return false;
          t.begin = (uint64_t)ps.num(); ps.eat(',');
          t.end = (uint64_t)ps.num();
          if (!ps.eat(']')) return false;
        } else ps.skip_value();
        ps.eat(',');
      }
      f.tensors.push_back(std::move(t));
    }
    ps.eat(',');
  }
  return ps.ok;
}

}  // namespace

ODB_API void* odb_st_open(const char* path) {
  File* f = new File();
  f->fd = ::open(path, O_RDONLY);
  if (f->fd < 0) { f->error = "cannot open file"; return f; }
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 8) { f->error = "file too small"; return f; }
  f->size = (size_t)st.st_size;
  f->map = (uint8_t*)mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
  if (f->map == MAP_FAILED) { f->map = nullptr; f->error = "mmap failed"; return f; }
  uint64_t hlen;
  memcpy(&hlen, f->map, 8);
  if (hlen > f->size - 8 || hlen > (1ull << 28)) { f->error = "implausible header length"; return f; }
  f->data_off = 8 + hlen;
  if (!parse_header(*f, (const char*)f->map + 8, (size_t)hlen)) { f->error = "malformed header"; return f; }
  for (auto& t : f->tensors)
    // overflow-safe: compare against the payload size instead of adding the (attacker-controlled) offset to data_off
    if (f->data_off > f->size || t.end < t.begin || t.end > f->size - f->data_off) { f->error = "tensor range outside file: " + t.name; break; }
  return f;
}
ODB_API const char* odb_st_error(void* h) { File* f = (File*)h; return f->error.empty() ? nullptr : f->error.c_str(); }
ODB_API int odb_st_count(void* h) { return (int)((File*)h)->tensors.size(); }
ODB_API const char* odb_st_name(void* h, int i) { return ((File*)h)->tensors[i].name.c_str(); }
ODB_API const char* odb_st_dtype(void* h, int i) { return ((File*)h)->tensors[i].dtype.c_str(); }
ODB_API int odb_st_ndim(void* h, int i) { return (int)((File*)h)->tensors[i].shape.size(); }
ODB_API int64_t odb_st_dim(void* h, int i, int d) { return ((File*)h)->tensors[i].shape[d]; }
ODB_API uint64_t odb_st_nbytes(void* h, int i) { auto& t = ((File*)h)->tensors[i]; return t.end - t.begin; }
ODB_API const void* odb_st_data(void* h, int i) { File* f = (File*)h; return f->map + f->data_off + f->tensors[i].begin; }
ODB_API void odb_st_close(void* h) {
  File* f = (File*)h;
  if (f->map) munmap(f->map, f->size);
  if (f->fd >= 0) ::close(f->fd);
  delete f;
}
