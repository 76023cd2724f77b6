// Host tokenizer of the C++ layer: the `tokenizer.json` pipeline the reference runs through the
// HF `tokenizers` crate inside model2vec-rs (encode_with_args -> tokenizer.encode_batch_fast(..,
// add_special_tokens = false), reference call site src/search/mod.rs:69), restated for the subset of
// components a SentencePiece-style static-embedding model uses:
//   normalizer     Sequence | Lowercase | Replace (string pattern, or the regex " {2,}") | Strip |
//                  Prepend | Precompiled (the SentencePiece charsmap carried by tokenizer.json --
//                  nmt_nfkc for the XLM-R family the reference's model uses: double-array trie +
//                  replacement blob, applied grapheme by grapheme exactly as tokenizers'
//                  normalizers/precompiled.rs does, quirks included; extended grapheme clusters per
//                  UAX #29 from generated property tables, grapheme_break.inc) |
//                  NFC/NFD/NFKC/NFKD/Nmt (identity on printable ASCII; a line with other characters
//                  is REFUSED with an error -- those composition tables are not restated here)
//   pre_tokenizer  Metaspace (replacement, prepend_scheme always|first|never, split) | WhitespaceSplit |
//                  Sequence of those
//   added_tokens   split out of the text before the model sees it, as AddedVocabulary::extract_and_normalize does:
//                  normalized = false tokens on the raw text, normalized = true tokens on the normalised
//                  pieces; leftmost-longest, lstrip / rstrip honoured; single_word = true is refused.
//                  (A document line that contains "<s>" or "<mask>" literally gets that token's id.)
//   model          Unigram (vocab [[token, score]...], unk_id, byte_fallback = false): Viterbi over a
//                  byte trie exactly as tokenizers' `encode_optimized` (f64 scores, strict > on ties,
//                  unknown characters at min_score - 10, consecutive unknowns fused into one token)
// Anything else in tokenizer.json fails the constructor with a clear message.  Output ids are
// checked token for token against HF `tokenizers` on synthetic vocabularies (tests/test_host_cpp.py);
// the real model's tokenizer.json is not available offline (parity unpinned for it).
#pragma once

#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "semtools_host.hpp"
#include "semtools_store.hpp"   // Json

namespace semtools {

class HfTokenizer : public Tokenizer {
 public:
  explicit HfTokenizer(const std::string &tokenizer_json_path);
  // ids of one line as encode_with_args prepares them: the id of the model's `unk_token` removed -- when the
  // model section NAMES one (WordPiece / BPE style "unk_token": "<tok>"; model2vec-rs reads exactly that key,
  // [UPSTREAM-MEMORY]).  A Unigram section carries `unk_id` instead, so -- as in the Python host and in
  // model2vec itself -- nothing is dropped and an unknown character pools the unk row.
  std::vector<uint32_t> encode(const std::string &text) const override;
  // ids exactly as tokenizer.encode(text, add_special_tokens = false).ids
  std::vector<uint32_t> encode_raw(const std::string &text) const;
  // the normalizer alone (tokenizer.normalizer.normalize_str)
  std::string normalize_str(const std::string &text) const { return normalize(text); }
  size_t median_token_length() const override { return median_len_; }
  size_t vocab_size() const { return tokens_.size(); }
  bool has_unk() const { return has_unk_; }
  uint32_t unk_id() const { return unk_id_; }
  bool drops_unk() const { return drop_unk_; }
  // fnv1a64 of the file: part of the store's model fingerprint
  uint64_t fingerprint() const { return file_hash_; }

 private:
  struct NormStep { int kind; std::string a, b; bool left = true, right = true; int map = -1; };
  // SentencePiece precompiled charsmap: darts-clone double array + NUL-separated replacement strings
  struct Charsmap {
    std::vector<uint32_t> trie;
    std::string normalized;
    // offset of the replacement of the SHORTEST key that is a prefix of p[0..n), or -1
    // (spm_precompiled: common_prefix_search(...)[0])
    int64_t first_prefix(const char *p, size_t n) const;
    bool ascii_plain[128] = {};   // printable ASCII bytes the map leaves alone when they stand alone
  };
  std::string apply_charsmap(const Charsmap &m, const std::string &s) const;
  struct PreStep { int kind; std::string replacement; int prepend = 0; bool split = true; };
  struct Added { std::string content; uint32_t id = 0; bool lstrip = false, rstrip = false; };
  struct Seg { int64_t id; std::string text; size_t start; };          // id < 0: plain text
  void split_added(const std::string &s, const std::vector<Added> &set, std::vector<Seg> &out) const;
  std::string normalize(const std::string &text) const;
  void pre_tokenize(const std::string &normalized, std::vector<std::string> &pieces, bool at_origin = true) const;
  void unigram(const std::string &piece, std::vector<uint32_t> &out) const;
  void add_norm(const Json &j);
  void add_pre(const Json &j);

  std::vector<NormStep> norm_;
  std::vector<Charsmap> maps_;
  std::vector<Added> added_raw_, added_norm_;                          // normalized = false / true
  std::vector<PreStep> pre_;
  std::vector<std::string> tokens_;
  std::vector<double> scores_;
  double min_score_ = 0.0;
  bool has_unk_ = false;           // Unigram unk_id: what unknown characters are mapped to
  uint32_t unk_id_ = 0;
  bool drop_unk_ = false;          // model.unk_token named and found: encode() removes drop_id_
  uint32_t drop_id_ = 0;
  size_t median_len_ = 1;
  uint64_t file_hash_ = 0;
  // byte trie, flattened after construction: node n's children are child_byte_/child_node_
  // [first_child_[n], first_child_[n + 1]) sorted by byte (the root has a direct 256-entry table);
  // terminal_[node] = token id or -1
  uint32_t step(uint32_t node, unsigned char c) const;      // child or UINT32_MAX
  std::vector<uint32_t> first_child_, child_node_;
  std::vector<uint8_t> child_byte_;
  uint32_t root_[256];
  std::vector<int32_t> terminal_;
};

// Extended grapheme clusters (UAX #29, GB1-GB13 including GB9c and GB11) of a UTF-8 string: the END offset
// of every cluster, in order (what unicode_segmentation::graphemes(true) yields).  Bytes that are not valid
// UTF-8 stand alone.
std::vector<size_t> grapheme_ends(const std::string &s);

}  // namespace semtools
