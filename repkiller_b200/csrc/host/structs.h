// Host-side types of the drop-in: same names and meaning as /root/reference/src/structs.h, written for this
// repo.  Differences on purpose: the 1-byte packing is scoped to FragFile only (the reference's file-scope
// `#pragma pack(1)`, structs.h:2, also packs every type included after it, which misaligns std::mutex —
// SURVEY.md fact 4), and the record's long double is kept as 16 opaque bytes.
#pragma once

#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

#pragma pack(push, 1)
struct FragFile {          // reference: structs.h:12-51, 109 bytes
  int64_t diag;            // xStart - yStart
  uint64_t xStart, yStart, xEnd, yEnd;
  uint64_t length, ident, score;
  float similarity;
  uint64_t seqX, seqY;
  int64_t block;
  char strand;             // 'f' forward; anything else is handled as reverse (commonFunctions.cpp:52-53)
  unsigned char evalue[16];
};
#pragma pack(pop)
static_assert(sizeof(FragFile) == 109, "record layout is part of the C ABI (RK_FRAG_BYTES)");

struct Sequence {          // reference: structs.h:54-58
  Sequence(uint64_t id, uint64_t len) : id(id), len(len) {}
  uint64_t id;
  uint64_t len;            // header value + 1 (FragmentsDatabase.cpp:62,65)
};

typedef std::vector<const FragFile *> FragsGroup;  // reference: structs.h:75-76
typedef std::vector<FragsGroup *> FGList;

class sequence_manager {   // reference: structs.h:79-91, class_structs.cpp
  std::string header;
 public:
  std::vector<Sequence> sequences;
  void read_header(const std::string &input_header) { header = input_header; }
  void write_header(std::ostream &out) const { out << header; }
  const std::string &raw_header() const { return header; }
  const Sequence &get_sequence_by_label(uint64_t label) const { return sequences[label]; }
  uint64_t get_maximum_length() const {
    uint64_t m = 0;
    for (const auto &s : sequences) m = s.len > m ? s.len : m;
    return m;
  }
  uint64_t get_number_of_sequences() const { return sequences.size(); }
};
