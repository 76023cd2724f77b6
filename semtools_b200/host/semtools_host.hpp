// C++ host side above the C ABI: the reference's search interface for this path
// (src/search/mod.rs: Document :18-22, SearchConfig :32-38, SearchResult :40-47,
// create_document_from_content :49-75, search_documents :77-120, search_files :122-143)
// and its rendering (src/cmds/search.rs:23-63, src/json_mode.rs:17-30), same names and
// argument meaning.  All arithmetic happens in libsemtools_b200.so; this layer keeps the
// row <-> (document, line) bookkeeping, builds context windows and prints.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/semtools_b200.h"

namespace semtools {

struct StbError : std::runtime_error {
  int status;
  StbError(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

struct Document {            // src/search/mod.rs:18-22 (embeddings live in HBM: rows [row_start, +lines.size()))
  std::string filename;
  std::vector<std::string> lines;
  uint64_t row_start = 0;
};

struct SearchConfig {        // src/search/mod.rs:32-38; CLI defaults src/bin/semtools.rs:61-74
  size_t n_lines = 3;
  size_t top_k = 3;
  std::optional<double> max_distance;
  bool ignore_case = false;
};

struct SearchResult {        // src/search/mod.rs:40-47
  std::string filename;
  std::vector<std::string> lines;
  size_t start = 0, end = 0, match_line = 0;
  double distance = 0.0;
};

// str::lines(): split on '\n', strip one trailing '\r', no trailing empty line (mod.rs:55)
std::vector<std::string> rust_lines(const std::string &content);
// ASCII-only stand-in for str::to_lowercase (mod.rs:64): non-ASCII bytes pass through
std::string to_lowercase_ascii(const std::string &s);
// `{}` of an f64: shortest round-trip digits, positional, "1" for 1.0
std::string rust_display_f64(double x);
// serde_json number: shortest round-trip; "1.0" for integral values; 1e-7 style exponents
std::string json_f64(double x);
std::string json_string(const std::string &s);

// Host tokenizer interface (tokenisation stays on the CPU; model2vec's own tokenizer is
// HF `tokenizers`).  Returns the ids of one line, already unk-dropped.
struct Tokenizer {
  virtual ~Tokenizer() = default;
  virtual std::vector<uint32_t> encode(const std::string &text) const = 0;
};
// WordLevel + whitespace split over a vocabulary file (one token per line, id = line
// number); unknown words are dropped (= unk removal of encode_with_args).
struct WordLevelTokenizer : Tokenizer {
  explicit WordLevelTokenizer(const std::string &vocab_path);
  std::vector<uint32_t> encode(const std::string &text) const override;
  std::vector<std::pair<std::string, uint32_t>> vocab;   // sorted by token
};

class Searcher {
 public:
  explicit Searcher(int device = 0);
  ~Searcher();
  Searcher(const Searcher &) = delete;
  // StaticModel tensors (src/cmds/search.rs:123-128)
  void load_table(const float *E, uint64_t V, bool normalize);
  // create_document_from_content (mod.rs:49-75): returns false for empty content (-> None)
  bool add_document(const std::string &filename, const std::string &content, const Tokenizer &tok, bool ignore_case);
  // same with embeddings computed elsewhere
  bool add_document_embeddings(const std::string &filename, const std::vector<std::string> &lines, const float *emb);
  // encode_single (mod.rs:138): 512-token truncation
  std::vector<float> encode_single(const std::string &query, const Tokenizer &tok) const;
  // search_documents (mod.rs:77-120)
  std::vector<SearchResult> search_documents(const std::vector<float> &query_embedding, const SearchConfig &cfg) const;
  uint64_t rows() const;

 private:
  stb_ctx *ctx_ = nullptr;
  stb_table *table_ = nullptr;
  stb_corpus *corpus_ = nullptr;
  std::vector<Document> docs_;
};

// print_search_results (src/cmds/search.rs:35-63) and SearchOutput JSON (:164-169)
std::string format_search_results(const std::vector<SearchResult> &results, bool is_tty);
std::string search_output_json(const std::vector<SearchResult> &results);

}  // namespace semtools
