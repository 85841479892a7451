// 10x MatrixMarket (.mtx) loader -> compact-index CSR with the reference's pandas semantics.
//
// Replaces load_matrix(), src/cell_type_training.py:9-17:
//     df = pd.read_csv(file, header=None, skiprows=3, delim_whitespace=True,
//                      names=['gene', 'barcode', 'p'])
//     df = df.pivot_table(index='barcode', columns='gene', values='p', fill_value=0)
// i.e. skip exactly 3 lines (the dims line is ignored), rows are the
// distinct barcode ids ascending, columns the distinct gene ids ascending (absent ids are
// dropped), duplicate (gene, barcode) entries are averaged in float64, explicit zeros keep
// their row/column alive, input order is irrelevant.
//
// Host-only C++ (no CUDA): the text parse is chunked over std::threads, the CSR is built with
// a counting sort over rows and a per-row sort over columns.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cellcomm_b200.h"

namespace cc {
void set_error(const char* fmt, ...);
}

struct cc_csr {
  int64_t rows = 0, cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colidx;
  std::vector<float> values;
  std::vector<double> values64;
  std::vector<int64_t> row_ids, col_ids;
};

namespace {

struct Triplets {
  std::vector<int64_t> gene, barcode;
  std::vector<double> val;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// parse one whitespace-separated field as a number; integers take the fast path
inline bool parse_number(const char*& p, const char* end, double* out, bool* is_int, int64_t* iv) {
  while (p < end && is_space(*p)) ++p;
  if (p >= end || *p == '\n') return false;
  const char* s = p;
  bool neg = false;
  if (*p == '-' || *p == '+') {
    neg = (*p == '-');
    ++p;
  }
  int64_t v = 0;
  int digits = 0;
  while (p < end && *p >= '0' && *p <= '9') {
    v = v * 10 + (*p - '0');
    ++p;
    ++digits;
  }
  if (digits > 0 && digits <= 18 && (p >= end || is_space(*p) || *p == '\n')) {
    *iv = neg ? -v : v;
    *out = (double)*iv;
    *is_int = true;
    return true;
  }
  // general float: find the field end and use strtod on a bounded copy
  const char* q = s;
  while (q < end && !is_space(*q) && *q != '\n') ++q;
  char buf[64];
  size_t len = (size_t)(q - s);
  if (len == 0 || len >= sizeof(buf)) return false;
  memcpy(buf, s, len);
  buf[len] = 0;
  char* e = nullptr;
  errno = 0;
  double d = strtod(buf, &e);
  if (e != buf + len) return false;
  *out = d;
  *is_int = false;
  *iv = (int64_t)d;
  p = q;
  return true;
}

// parse lines in [p, end); returns false (with line text) on a malformed line
bool parse_chunk(const char* p, const char* end, Triplets* out, std::string* err) {
  while (p < end) {
    // skip blank lines (pandas skip_blank_lines=True)
    const char* ls = p;
    while (p < end && is_space(*p)) ++p;
    if (p >= end) break;
    if (*p == '\n') {
      ++p;
      continue;
    }
    double g, b, v;
    bool gi, bi, vi;
    int64_t giv, biv, viv;
    bool ok = parse_number(p, end, &g, &gi, &giv) && parse_number(p, end, &b, &bi, &biv) &&
              parse_number(p, end, &v, &vi, &viv) && gi && bi;
    if (ok) {
      while (p < end && is_space(*p)) ++p;
      ok = (p >= end || *p == '\n');
    }
    if (!ok) {
      const char* le = ls;
      while (le < end && *le != '\n') ++le;
      *err = std::string(ls, std::min<size_t>((size_t)(le - ls), 80));
      return false;
    }
    out->gene.push_back(giv);
    out->barcode.push_back(biv);
    out->val.push_back(v);
    if (p < end) ++p;  // newline
  }
  return true;
}

// ids -> (sorted distinct ids, compact index of every element)
void compact_ids(const int64_t* ids, int64_t n, std::vector<int64_t>* distinct,
                 std::vector<int32_t>* index) {
  index->resize((size_t)n);
  distinct->clear();
  if (n == 0) return;
  int64_t lo = ids[0], hi = ids[0];
  for (int64_t i = 1; i < n; ++i) {
    lo = std::min(lo, ids[i]);
    hi = std::max(hi, ids[i]);
  }
  const int64_t range = hi - lo + 1;
  if (range <= std::max<int64_t>(4 * n, 1 << 20)) {
    std::vector<int32_t> lut((size_t)range, -1);
    for (int64_t i = 0; i < n; ++i) lut[(size_t)(ids[i] - lo)] = 0;
    int32_t next = 0;
    for (int64_t k = 0; k < range; ++k)
      if (lut[(size_t)k] == 0) {
        lut[(size_t)k] = next++;
        distinct->push_back(lo + k);
      }
    for (int64_t i = 0; i < n; ++i) (*index)[(size_t)i] = lut[(size_t)(ids[i] - lo)];
  } else {
    std::vector<int64_t> s(ids, ids + n);
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    *distinct = s;
    for (int64_t i = 0; i < n; ++i)
      (*index)[(size_t)i] =
          (int32_t)(std::lower_bound(s.begin(), s.end(), ids[i]) - s.begin());
  }
}

int build_csr(const int64_t* gene, const int64_t* barcode, const double* val, int64_t nnz,
              cc_csr** out) {
  cc_csr* c = new cc_csr();
  std::vector<int32_t> ri, ci;
  compact_ids(barcode, nnz, &c->row_ids, &ri);
  compact_ids(gene, nnz, &c->col_ids, &ci);
  c->rows = (int64_t)c->row_ids.size();
  c->cols = (int64_t)c->col_ids.size();
  // counting sort by row
  std::vector<int64_t> start((size_t)c->rows + 1, 0);
  for (int64_t i = 0; i < nnz; ++i) ++start[(size_t)ri[(size_t)i] + 1];
  for (int64_t r = 0; r < c->rows; ++r) start[(size_t)r + 1] += start[(size_t)r];
  std::vector<int64_t> pos(start.begin(), start.end() - 1);
  struct Ent {
    int32_t col;
    double v;
  };
  std::vector<Ent> ents((size_t)nnz);
  for (int64_t i = 0; i < nnz; ++i) {
    int64_t& p = pos[(size_t)ri[(size_t)i]];
    ents[(size_t)p++] = Ent{ci[(size_t)i], val[(size_t)i]};
  }
  // per-row: sort by column (stable, so duplicates keep file order), average duplicates
  c->rowptr.assign((size_t)c->rows + 1, 0);
  c->colidx.reserve((size_t)nnz);
  c->values64.reserve((size_t)nnz);
  for (int64_t r = 0; r < c->rows; ++r) {
    Ent* b = ents.data() + start[(size_t)r];
    Ent* e = ents.data() + start[(size_t)r + 1];
    std::stable_sort(b, e, [](const Ent& x, const Ent& y) { return x.col < y.col; });
    for (Ent* q = b; q < e;) {
      Ent* g = q;
      double sum = 0.0;
      int64_t cnt = 0;
      while (g < e && g->col == q->col) {
        sum += g->v;
        ++cnt;
        ++g;
      }
      c->colidx.push_back(q->col);
      c->values64.push_back(cnt == 1 ? sum : sum / (double)cnt);
      q = g;
    }
    c->rowptr[(size_t)r + 1] = (int64_t)c->colidx.size();
  }
  c->values.resize(c->values64.size());
  for (size_t i = 0; i < c->values64.size(); ++i) c->values[i] = (float)c->values64[i];
  *out = c;
  return 0;
}

}  // namespace

extern "C" int cc_coo_to_csr(const int64_t* gene, const int64_t* barcode, const double* val,
                             int64_t nnz, cc_csr** out) {
  if (out == nullptr || nnz < 0) {
    cc::set_error("cc_coo_to_csr: bad arguments");
    return -1;
  }
  return build_csr(gene, barcode, val, nnz, out);
}

extern "C" int cc_mtx_load_csr(const char* path, cc_csr** out) {
  if (path == nullptr || out == nullptr) {
    cc::set_error("cc_mtx_load_csr: bad arguments");
    return -1;
  }
  int fd = open(path, O_RDONLY);
  if (fd < 0) {
    cc::set_error("cc_mtx_load_csr: cannot open %s: %s", path, strerror(errno));
    return -1;
  }
  struct stat sb;
  if (fstat(fd, &sb) != 0) {
    close(fd);
    cc::set_error("cc_mtx_load_csr: fstat failed on %s", path);
    return -1;
  }
  const size_t size = (size_t)sb.st_size;
  const char* data = nullptr;
  if (size > 0) {
    data = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (data == MAP_FAILED) {
      close(fd);
      cc::set_error("cc_mtx_load_csr: mmap failed on %s", path);
      return -1;
    }
  }
  // skip exactly 3 lines (skiprows=3): banner, comment, dims
  const char* p = data;
  const char* end = data + size;
  for (int l = 0; l < 3 && p < end; ++l) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    p = nl ? nl + 1 : end;
  }
  // chunk the body at line boundaries
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  const size_t body = (size_t)(end - p);
  if (body < (1u << 20)) nt = 1;
  std::vector<const char*> cut(nt + 1);
  cut[0] = p;
  cut[nt] = end;
  for (unsigned t = 1; t < nt; ++t) {
    const char* q = p + body / nt * t;
    if (q < cut[t - 1]) q = cut[t - 1];
    const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
    cut[t] = nl ? nl + 1 : end;
  }
  std::vector<Triplets> parts(nt);
  std::vector<std::string> errs(nt);
  std::vector<char> oks(nt, 1);
  std::vector<std::thread> threads;
  for (unsigned t = 0; t < nt; ++t)
    threads.emplace_back([&, t] { oks[t] = parse_chunk(cut[t], cut[t + 1], &parts[t], &errs[t]); });
  for (auto& th : threads) th.join();
  if (data) munmap((void*)data, size);
  close(fd);
  for (unsigned t = 0; t < nt; ++t)
    if (!oks[t]) {
      cc::set_error("cc_mtx_load_csr: malformed line in %s: '%s'", path, errs[t].c_str());
      return -1;
    }
  size_t nnz = 0;
  for (auto& q : parts) nnz += q.val.size();
  Triplets all;
  all.gene.reserve(nnz);
  all.barcode.reserve(nnz);
  all.val.reserve(nnz);
  for (auto& q : parts) {
    all.gene.insert(all.gene.end(), q.gene.begin(), q.gene.end());
    all.barcode.insert(all.barcode.end(), q.barcode.begin(), q.barcode.end());
    all.val.insert(all.val.end(), q.val.begin(), q.val.end());
    Triplets().gene.swap(q.gene);
    Triplets().barcode.swap(q.barcode);
    Triplets().val.swap(q.val);
  }
  return build_csr(all.gene.data(), all.barcode.data(), all.val.data(), (int64_t)nnz, out);
}

extern "C" void cc_csr_destroy(cc_csr* csr) { delete csr; }
extern "C" int64_t cc_csr_rows(const cc_csr* c) { return c->rows; }
extern "C" int64_t cc_csr_cols(const cc_csr* c) { return c->cols; }
extern "C" int64_t cc_csr_nnz(const cc_csr* c) { return (int64_t)c->colidx.size(); }
extern "C" const int64_t* cc_csr_rowptr(const cc_csr* c) { return c->rowptr.data(); }
extern "C" const int32_t* cc_csr_colidx(const cc_csr* c) { return c->colidx.data(); }
extern "C" const float* cc_csr_values(const cc_csr* c) { return c->values.data(); }
extern "C" const double* cc_csr_values64(const cc_csr* c) { return c->values64.data(); }
extern "C" const int64_t* cc_csr_row_ids(const cc_csr* c) { return c->row_ids.data(); }
extern "C" const int64_t* cc_csr_col_ids(const cc_csr* c) { return c->col_ids.data(); }
