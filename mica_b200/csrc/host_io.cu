// Host-side text munging of the file formats around the hot path: the fixed-column
// PDB reader that stands where Bio.PDB.PDBParser stands in the reference (utils/preprocessing.py:269,
// 275-298).  No device code: it lives in the library because a 160 k-atom docked model costs ~40 ms to
// parse with NumPy, ~20 ms here on one host thread and a few ms with one piece of the text per host thread, and the
// drop-in's end-to-end time is made of such milliseconds.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace {

// value of a numeric PDB field exactly as Python's float(text) gives it.  Fast path (Clinger): a decimal
// with <= 15 significant digits and <= 22 fraction digits is mantissa / 10^k with both exactly
// representable, so one IEEE division is the correctly rounded result.  Anything else goes to strtod.
bool parse_field(const char* p, int width, double* out) {
  int i = 0, e = width;
  while (i < e && (p[i] == ' ' || p[i] == '\t')) ++i;
  while (e > i && (p[e - 1] == ' ' || p[e - 1] == '\t' || p[e - 1] == '\r')) --e;
  if (i >= e) return false;
  bool neg = false;
  int j = i;
  if (p[j] == '-' || p[j] == '+') {
    neg = p[j] == '-';
    ++j;
  }
  unsigned long long mant = 0;
  int digits = 0, frac = 0;
  bool seen_point = false, simple = j < e;
  for (int k = j; k < e; ++k) {
    const char c = p[k];
    if (c >= '0' && c <= '9') {
      mant = mant * 10ull + (unsigned long long)(c - '0');
      ++digits;
      if (seen_point) ++frac;
    } else if (c == '.' && !seen_point) {
      seen_point = true;
    } else {
      simple = false;
      break;
    }
  }
  if (simple && digits > 0 && digits <= 15 && frac <= 22) {
    static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const double v = (double)mant / p10[frac];
    *out = neg ? -v : v;
    return true;
  }
  char buf[64];
  const int n = e - i;
  if (n <= 0 || n >= (int)sizeof(buf)) return false;
  memcpy(buf, p + i, (size_t)n);
  buf[n] = 0;
  char* end = nullptr;
  const double v = strtod(buf, &end);
  if (end != buf + n) return false;
  *out = v;
  return true;
}

// channel of an atom name / residue name (utils/preprocessing.py:254-263): stripped text compared
int bb_code(const char* f4) {
  if (f4[0] == ' ' && f4[3] == ' ' && f4[2] == ' ') {      // " X  ": the usual spelling of N, C, O
    if (f4[1] == 'N') return 1;
    if (f4[1] == 'C') return 2;
    if (f4[1] == 'O') return 3;
    return -1;
  }
  char t[5];
  int n = 0;
  for (int i = 0; i < 4; ++i)
    if (f4[i] != ' ') t[n++] = f4[i];
  t[n] = 0;
  if (!strcmp(t, "CA")) return 0;
  if (!strcmp(t, "N")) return 1;
  if (!strcmp(t, "C")) return 2;
  if (!strcmp(t, "O")) return 3;
  return -1;
}
int aa_code(const char* f3) {
  if (f3[0] == ' ' || f3[1] == ' ' || f3[2] == ' ') return -1;   // the 20 names fill the field
  const unsigned k = ((unsigned)(unsigned char)f3[0] << 16) | ((unsigned)(unsigned char)f3[1] << 8) | (unsigned char)f3[2];
#define AA3(a, b, c) ((unsigned)(a) << 16 | (unsigned)(b) << 8 | (unsigned)(c))
  switch (k) {
    case AA3('A', 'L', 'A'): return 4;
    case AA3('C', 'Y', 'S'): return 5;
    case AA3('A', 'S', 'P'): return 6;
    case AA3('G', 'L', 'U'): return 7;
    case AA3('P', 'H', 'E'): return 8;
    case AA3('G', 'L', 'Y'): return 9;
    case AA3('H', 'I', 'S'): return 10;
    case AA3('I', 'L', 'E'): return 11;
    case AA3('L', 'Y', 'S'): return 12;
    case AA3('L', 'E', 'U'): return 13;
    case AA3('M', 'E', 'T'): return 14;
    case AA3('A', 'S', 'N'): return 15;
    case AA3('P', 'R', 'O'): return 16;
    case AA3('G', 'L', 'N'): return 17;
    case AA3('A', 'R', 'G'): return 18;
    case AA3('S', 'E', 'R'): return 19;
    case AA3('T', 'H', 'R'): return 20;
    case AA3('V', 'A', 'L'): return 21;
    case AA3('T', 'R', 'P'): return 22;
    case AA3('T', 'Y', 'R'): return 23;
    default: return -1;
  }
#undef AA3
}

// ---- one contiguous piece of the text (whole lines), parsed by one host thread
struct PdbOut {
  float* xyz;
  uint8_t* fields;
  float* occupancy;
  int32_t* model;
  int8_t* bb_ch;
  int8_t* aa_ch;
  bool want_dup;
};

struct PdbPiece {
  const char* begin;
  const char* end;
  // counting pass
  int64_t n_records = 0;
  int32_t n_models = 0;
  // parsing pass: records [n0, n0 + n_records) and MODEL count m0 before the piece
  int64_t n0 = 0;
  int32_t m0 = 0;
  int64_t n_res = 0;            // residue runs that START in this piece (its first record always starts one)
  int64_t n_written = 0;
  bool dup = false;
  int64_t err_record = -1;      // first record whose coordinates cannot be parsed
  int err_axis = 0;
  // the first and the last residue run of the piece, for stitching runs across piece boundaries
  unsigned char first_key[12], last_key[12];
  unsigned first_names[256], last_names[256];
  int n_first_names = 0, n_last_names = 0;
  bool single_run = true;       // the piece holds one residue run only (first run == last run)
};

inline bool is_record(const char* p, int64_t len, int with_hetatm) {
  return len >= 6 && (memcmp(p, "ATOM  ", 6) == 0 || (with_hetatm && memcmp(p, "HETATM", 6) == 0));
}

void count_piece(PdbPiece* pc, int with_hetatm) {
  const char* p = pc->begin;
  const char* const end = pc->end;
  int64_t n = 0;
  int32_t m = 0;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const int64_t len = le - p;
    if (len >= 5 && memcmp(p, "MODEL", 5) == 0)
      ++m;
    else if (is_record(p, len, with_hetatm))
      ++n;
    if (!nl) break;
    p = nl + 1;
  }
  pc->n_records = n;
  pc->n_models = m;
}

void parse_piece(PdbPiece* pc, int with_hetatm, int64_t capacity, const PdbOut& o) {
  int64_t n = pc->n0;
  int32_t n_model = pc->m0;
  unsigned names[256];
  int n_names = 0;
  unsigned char last_res[12];
  bool have_res = false, in_first = true;
  const char* p = pc->begin;
  const char* const end = pc->end;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* le = nl ? nl : end;
    const int64_t len = le - p;
    if (len >= 5 && memcmp(p, "MODEL", 5) == 0) {
      ++n_model;
    } else if (is_record(p, len, with_hetatm)) {
      if (n < capacity) {
        // a line of 61+ characters is read in place (a '\r' can only be its last character); shorter ones are
        // padded with blanks to the 60 columns the fields below address
        char pad[61];
        const char* line = p;
        if (len < 61) {
          const int m = len < 60 ? (int)len : 60;
          memcpy(pad, p, (size_t)m);
          for (int i = m; i < 60; ++i) pad[i] = ' ';
          for (int i = 0; i < 60; ++i)
            if (pad[i] == '\r') pad[i] = ' ';
          line = pad;
        }
        double v[3];
        for (int a = 0; a < 3; ++a)
          if (!parse_field(line + 30 + 8 * a, 8, &v[a])) {
            pc->err_record = n;
            pc->err_axis = a;
            return;
          }
        o.xyz[3 * n + 0] = (float)v[0];
        o.xyz[3 * n + 1] = (float)v[1];
        o.xyz[3 * n + 2] = (float)v[2];
        uint8_t* f = o.fields + 16 * n;
        memcpy(f + 0, line + 12, 4);
        f[4] = (uint8_t)line[16];
        memcpy(f + 5, line + 17, 3);
        f[8] = (uint8_t)line[21];
        memcpy(f + 9, line + 22, 5);
        f[14] = (p[0] == 'H') ? 1 : 0;
        f[15] = 0;
        double occ = 0.0;
        if (o.occupancy) o.occupancy[n] = parse_field(line + 54, 6, &occ) ? (float)occ : 0.f;
        if (o.model) o.model[n] = n_model;
        if (o.bb_ch) o.bb_ch[n] = (int8_t)bb_code(line + 12);
        if (o.aa_ch) o.aa_ch[n] = (int8_t)aa_code(line + 17);
        unsigned char res[12];
        memcpy(res, f + 8, 7);               // chain, resSeq + iCode, HETATM flag
        memcpy(res + 7, &n_model, 4);
        res[11] = 0;
        if (!have_res || memcmp(res, last_res, 12) != 0) {
          if (have_res) {
            if (in_first) {                  // the first run of the piece ends here: keep its names
              memcpy(pc->first_names, names, sizeof(unsigned) * (size_t)n_names);
              pc->n_first_names = n_names;
              in_first = false;
              pc->single_run = false;
            }
          } else {
            memcpy(pc->first_key, res, 12);
          }
          ++pc->n_res;
          memcpy(last_res, res, 12);
          have_res = true;
          n_names = 0;
        }
        if (o.want_dup) {
          unsigned nm;
          memcpy(&nm, f, 4);
          if (!pc->dup)
            for (int i = 0; i < n_names; ++i)
              if (names[i] == nm) pc->dup = true;
          if (n_names < 256) names[n_names++] = nm;
        }
        ++pc->n_written;
      }
      ++n;
    }
    if (!nl) break;
    p = nl + 1;
  }
  if (have_res) {
    memcpy(pc->last_key, last_res, 12);
    memcpy(pc->last_names, names, sizeof(unsigned) * (size_t)n_names);
    pc->n_last_names = n_names;
    if (in_first) {
      memcpy(pc->first_names, names, sizeof(unsigned) * (size_t)n_names);
      pc->n_first_names = n_names;
    }
  }
}

int pdb_threads(int64_t nbytes) {
  const char* env = getenv("MICA_PDB_THREADS");
  int t = env ? atoi(env) : 0;
  if (t <= 0) {
    t = (int)std::thread::hardware_concurrency();
    if (t > 16) t = 16;
  }
  if (t > 64) t = 64;
  int64_t min_piece = 256 << 10;                     // a piece below 256 KB is not worth a thread
  if (const char* mp = getenv("MICA_PDB_MIN_PIECE")) {   // (tests cut small texts into many pieces)
    const long long v = atoll(mp);
    if (v > 0) min_piece = v;
  }
  const int64_t by_size = nbytes / min_piece;
  if (t > by_size) t = (int)by_size;
  return t < 1 ? 1 : t;
}

template <typename F>
void run_pieces(std::vector<PdbPiece>& pieces, F f) {
  if (pieces.size() == 1) {
    f(&pieces[0]);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(pieces.size() - 1);
  for (size_t i = 1; i < pieces.size(); ++i) pool.emplace_back(f, &pieces[i]);
  f(&pieces[0]);
  for (auto& t : pool) t.join();
}

}  // namespace

// Parses the ATOM (and, with_hetatm != 0, HETATM) records of a PDB text.  Per record r:
//   xyz[3 r ..]      columns 31-38 / 39-46 / 47-54 as float32 (float(text) rounded to float32, as Bio.PDB stores them)
//   fields[16 r ..]  0-3 atom name (cols 13-16, with its spacing), 4 altloc (col 17), 5-7 residue name
//                    (cols 18-20), 8 chain (col 22), 9-13 resSeq + iCode (cols 23-27), 14 = 1 for HETATM, 15 = 0
//   occupancy[r]     columns 55-60 (0 when blank), model[r] = number of MODEL records seen before it
//   bb_ch[r] / aa_ch[r]  (nullable) backbone channel 0..3 | -1 and amino-acid channel 4..23 | -1
//   info                 (nullable) [0] residues (runs of equal model / chain / record type / resSeq+iCode),
//                        [1] = 1 when two records share (model, chain, record type, resSeq+iCode, atom name) --
//                        alternate locations or a name defined twice: the caller then applies Bio.PDB's rule
// Returns the number of records (at most `capacity` are written; call with capacity 0 to count), or a
// negative MICA_ERR_* when a coordinate field cannot be parsed.
//
// The text is cut into pieces of whole lines, one host thread each (MICA_PDB_THREADS, default = the hardware
// threads, at most 16): a counting pass gives every piece its first record index and MODEL count, the
// parsing pass fills disjoint output ranges, and the residue runs / duplicate names that straddle a cut are
// stitched from the pieces' first and last runs.  The result does not depend on the number of pieces.
extern "C" int64_t mica_parse_pdb(const char* text, int64_t nbytes, int with_hetatm, int64_t capacity, float* xyz,
                                  uint8_t* fields, float* occupancy, int32_t* model, int8_t* bb_ch, int8_t* aa_ch,
                                  int64_t* info) {
  if (!text || nbytes < 0) return mica::set_error(MICA_ERR_INVALID, "null PDB text");
  const int T = pdb_threads(nbytes);
  std::vector<PdbPiece> pieces((size_t)T);
  const char* const end = text + nbytes;
  {
    const char* b = text;
    for (int c = 0; c < T; ++c) {
      const char* e = end;
      if (c + 1 < T) {
        e = text + nbytes * (c + 1) / T;
        if (e < b) e = b;
        const char* nl = e < end ? (const char*)memchr(e, '\n', (size_t)(end - e)) : nullptr;
        e = nl ? nl + 1 : end;
      }
      pieces[(size_t)c].begin = b;
      pieces[(size_t)c].end = e;
      b = e;
    }
  }
  run_pieces(pieces, [&](PdbPiece* pc) { count_piece(pc, with_hetatm); });
  int64_t n = 0;
  int32_t m = 0;
  for (auto& pc : pieces) {
    pc.n0 = n;
    pc.m0 = m;
    n += pc.n_records;
    m += pc.n_models;
  }
  if (capacity > 0) {
    const PdbOut o{xyz, fields, occupancy, model, bb_ch, aa_ch, info != nullptr};
    run_pieces(pieces, [&](PdbPiece* pc) { parse_piece(pc, with_hetatm, capacity, o); });
    for (auto& pc : pieces)
      if (pc.err_record >= 0)
        return mica::set_error(MICA_ERR_INVALID, "PDB record %lld: cannot parse coordinate %d",
                               (long long)pc.err_record, pc.err_axis);
  }
  if (info) {
    int64_t n_res = 0;
    bool dup = false;
    // the residue run open at the end of the pieces seen so far
    unsigned char tail_key[12];
    std::vector<unsigned> tail_names;
    bool have_tail = false;
    for (auto& pc : pieces) {
      if (pc.n_written == 0) continue;
      n_res += pc.n_res;
      dup = dup || pc.dup;
      const bool joins = have_tail && memcmp(tail_key, pc.first_key, 12) == 0;
      if (joins) {
        --n_res;                             // the piece's first run continues the open one
        for (int i = 0; i < pc.n_first_names && !dup; ++i)
          for (unsigned nm : tail_names)
            if (nm == pc.first_names[i]) {
              dup = true;
              break;
            }
      }
      if (joins && pc.single_run) {
        tail_names.insert(tail_names.end(), pc.first_names, pc.first_names + pc.n_first_names);
      } else {
        memcpy(tail_key, pc.last_key, 12);
        tail_names.assign(pc.last_names, pc.last_names + pc.n_last_names);
      }
      have_tail = true;
    }
    info[0] = n_res;
    info[1] = dup ? 1 : 0;
  }
  return n;
}
