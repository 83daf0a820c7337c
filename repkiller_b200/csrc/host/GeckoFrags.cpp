#include "GeckoFrags.h"

#include <cstring>

namespace {
inline uint64_t load_be64(const unsigned char *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return __builtin_bswap64(v);
}
inline void store_be64(unsigned char *p, uint64_t v) {
  v = __builtin_bswap64(v);
  memcpy(p, &v, 8);
}
// field offsets inside a record: structs.h:12-51 under pack(1)
constexpr size_t O_DIAG = 0, O_XSTART = 8, O_YSTART = 16, O_XEND = 24, O_YEND = 32, O_LENGTH = 40, O_IDENT = 48, O_SCORE = 56,
                 O_SIM = 64, O_SEQX = 68, O_SEQY = 76, O_BLOCK = 84, O_STRAND = 92, O_EVALUE = 93;
static_assert(O_EVALUE + 16 == sizeof(FragFile), "record layout");
}  // namespace

void gecko_frags_lengths(const unsigned char *file, uint64_t *seqx_len, uint64_t *seqy_len) {
  *seqx_len = load_be64(file);
  *seqy_len = load_be64(file + 8);
}

void gecko_frags_decode(const unsigned char *file, uint64_t first, uint64_t n, FragFile *out) {
  const unsigned char *p = file + GECKO_FRAGS_HEADER_BYTES + first * sizeof(FragFile);
  for (uint64_t i = 0; i < n; ++i, p += sizeof(FragFile)) {
    FragFile &f = out[i];
    f.xStart = load_be64(p + O_XSTART);
    f.yStart = load_be64(p + O_YSTART);
    f.diag = (int64_t)f.xStart - (int64_t)f.yStart;  // FragmentsDatabase.cpp:32
    f.xEnd = load_be64(p + O_XEND);
    f.yEnd = load_be64(p + O_YEND);
    f.length = load_be64(p + O_LENGTH);
    f.score = load_be64(p + O_SCORE);
    uint32_t sim_bits;
    memcpy(&sim_bits, p + O_SIM, 4);
    sim_bits = __builtin_bswap32(sim_bits);
    memcpy(&f.similarity, &sim_bits, 4);
    f.ident = (uint64_t)f.similarity;                // :39 — the reference takes ident from the similarity column
    f.seqX = 0, f.seqY = 1;                          // :41-42
    f.block = (int64_t)load_be64(p + O_BLOCK);
    f.strand = (char)p[O_STRAND];
    memset(f.evalue, 0, sizeof f.evalue);            // :43
  }
}

void gecko_frags_encode(uint64_t seqx_len, uint64_t seqy_len, const FragFile *rec, uint64_t n, unsigned char *file) {
  store_be64(file, seqx_len);
  store_be64(file + 8, seqy_len);
  unsigned char *p = file + GECKO_FRAGS_HEADER_BYTES;
  for (uint64_t i = 0; i < n; ++i, p += sizeof(FragFile)) {
    const FragFile &f = rec[i];
    store_be64(p + O_DIAG, (uint64_t)f.diag);
    store_be64(p + O_XSTART, f.xStart);
    store_be64(p + O_YSTART, f.yStart);
    store_be64(p + O_XEND, f.xEnd);
    store_be64(p + O_YEND, f.yEnd);
    store_be64(p + O_LENGTH, f.length);
    store_be64(p + O_IDENT, f.ident);
    store_be64(p + O_SCORE, f.score);
    uint32_t sim_bits;
    memcpy(&sim_bits, &f.similarity, 4);
    sim_bits = __builtin_bswap32(sim_bits);
    memcpy(p + O_SIM, &sim_bits, 4);
    store_be64(p + O_SEQX, f.seqX);
    store_be64(p + O_SEQY, f.seqY);
    store_be64(p + O_BLOCK, (uint64_t)f.block);
    p[O_STRAND] = (unsigned char)f.strand;
    for (int b = 0; b < 16; ++b) p[O_EVALUE + b] = f.evalue[15 - b];
  }
}
