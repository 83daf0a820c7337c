// GECKO's binary fragment container (".frags"), the file GECKO's own tools exchange before anything is printed as CSV
// (SURVEY.md §8f N4).  The reference repository reads the CSV only; the one trace of the binary form in it is the dead
// endianessConversion() at /root/reference/src/FragmentsDatabase.cpp:12-14.  The layout below restates GECKO's published
// writer (writeSequenceLength / writeFragment: every value most significant byte first) over the record the reference
// declares at /root/reference/src/structs.h:12-51 — PARITY UNPINNED: no GECKO binary and no reference reader exist here.
//
//   bytes 0..7    length of sequence X, uint64 big-endian      bytes 8..15   length of sequence Y
//   then one 109-byte record per fragment: the fields of struct FragFile in declaration order (diag, xStart, yStart,
//   xEnd, yEnd, length, ident, score, similarity, seqX, seqY, block, strand, evalue), each field byte-reversed.
//
// What a record loads as is defined through the CSV route, the only one the reference has: the FragFile readFragment
// (/root/reference/src/FragmentsDatabase.cpp:30-43) produces from a row that prints the record's values exactly —
// coordinates, length, score, block, strand and similarity as stored; ident := (uint64_t)similarity (:39),
// diag := xStart - yStart (:32), seqX := 0, seqY := 1, evalue := 0 (:41-43).  A .frags file and its CSV rendering therefore
// give the same database, the same groups and the same output lines.
#pragma once

#include <cstddef>
#include <cstdint>

#include "structs.h"

constexpr size_t GECKO_FRAGS_HEADER_BYTES = 16;

// does `bytes` have the size of a header plus a whole number of records?
inline bool gecko_frags_layout_ok(size_t bytes) {
  return bytes >= GECKO_FRAGS_HEADER_BYTES && (bytes - GECKO_FRAGS_HEADER_BYTES) % sizeof(FragFile) == 0;
}
inline uint64_t gecko_frags_count(size_t bytes) { return (bytes - GECKO_FRAGS_HEADER_BYTES) / sizeof(FragFile); }
// the two sequence lengths of the header
void gecko_frags_lengths(const unsigned char *file, uint64_t *seqx_len, uint64_t *seqy_len);
// records [first, first + n) of the file -> out[0..n)
void gecko_frags_decode(const unsigned char *file, uint64_t first, uint64_t n, FragFile *out);
// the inverse, for tools and tests: header and records as GECKO writes them (every stored field of `rec` verbatim)
void gecko_frags_encode(uint64_t seqx_len, uint64_t seqy_len, const FragFile *rec, uint64_t n, unsigned char *file);
