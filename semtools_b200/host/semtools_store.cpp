#include "semtools_store.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace semtools {

// ------------------------------------------------------------------ tiny JSON ----------
namespace {
struct P {
  const std::string &s;
  size_t i = 0;
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
  [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("json: ") + m); }
  Json value() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    Json v;
    char c = s[i];
    if (c == '{') {
      v.type = Json::Obj; ++i; ws();
      if (s[i] == '}') { ++i; return v; }
      for (;;) {
        ws(); Json k = value(); if (k.type != Json::Str) fail("key");
        ws(); if (s[i++] != ':') fail("colon");
        v.obj.emplace_back(k.str, value());
        ws(); if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; return v; }
        fail("object");
      }
    }
    if (c == '[') {
      v.type = Json::Arr; ++i; ws();
      if (s[i] == ']') { ++i; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws(); if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; return v; }
        fail("array");
      }
    }
    if (c == '"') {
      v.type = Json::Str; ++i;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\') {
          char e = s[++i];
          switch (e) {
            case 'n': v.str += '\n'; break; case 't': v.str += '\t'; break; case 'r': v.str += '\r'; break;
            case 'b': v.str += '\b'; break; case 'f': v.str += '\f'; break; case '/': v.str += '/'; break;
            case '\\': v.str += '\\'; break; case '"': v.str += '"'; break;
            case 'u': {
              unsigned cp = (unsigned)std::stoul(s.substr(i + 1, 4), nullptr, 16); i += 4;
              if (cp >= 0xD800 && cp < 0xDC00 && s.compare(i + 1, 2, "\\u") == 0) {   // surrogate pair
                unsigned lo = (unsigned)std::stoul(s.substr(i + 3, 4), nullptr, 16); i += 6;
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
              }
              if (cp < 0x80) v.str += (char)cp;
              else if (cp < 0x800) { v.str += (char)(0xC0 | (cp >> 6)); v.str += (char)(0x80 | (cp & 0x3F)); }
              else if (cp < 0x10000) { v.str += (char)(0xE0 | (cp >> 12)); v.str += (char)(0x80 | ((cp >> 6) & 0x3F)); v.str += (char)(0x80 | (cp & 0x3F)); }
              else { v.str += (char)(0xF0 | (cp >> 18)); v.str += (char)(0x80 | ((cp >> 12) & 0x3F)); v.str += (char)(0x80 | ((cp >> 6) & 0x3F)); v.str += (char)(0x80 | (cp & 0x3F)); }
              break;
            }
            default: fail("escape");
          }
          ++i;
        } else v.str += s[i++];
      }
      if (i >= s.size()) fail("string");
      ++i;
      return v;
    }
    if (s.compare(i, 4, "true") == 0) { v.type = Json::Bool; v.b = true; i += 4; return v; }
    if (s.compare(i, 5, "false") == 0) { v.type = Json::Bool; i += 5; return v; }
    if (s.compare(i, 4, "null") == 0) { i += 4; return v; }
    size_t j = i;
    while (j < s.size() && (std::isdigit((unsigned char)s[j]) || s[j] == '-' || s[j] == '+' || s[j] == '.' || s[j] == 'e' || s[j] == 'E')) ++j;
    if (j == i) fail("value");
    v.type = Json::Num; v.raw_num = s.substr(i, j - i); v.num = std::strtod(v.raw_num.c_str(), nullptr); i = j;
    return v;
  }
};
}  // namespace

const Json *Json::get(const std::string &key) const {
  for (const auto &kv : obj) if (kv.first == key) return &kv.second;
  return nullptr;
}
Json Json::parse(const std::string &text) { P p{text}; Json v = p.value(); return v; }

static std::string read_file(const std::string &p) {
  std::ifstream f(p, std::ios::binary);
  if (!f) throw std::runtime_error("cannot read " + p);
  std::stringstream ss; ss << f.rdbuf(); return ss.str();
}
static bool exists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
static void mkdirs(const std::string &p) {
  std::string cur;
  for (size_t i = 0; i <= p.size(); ++i) {
    if (i == p.size() || p[i] == '/') { if (!cur.empty()) ::mkdir(cur.c_str(), 0755); }
    if (i < p.size()) cur += p[i];
  }
}
static std::string home() {
  const char *h = std::getenv("HOME");
  if (!h || !*h) throw std::runtime_error("No home dir found?");
  return h;
}

// ------------------------------------------------------------------ Workspace (mod.rs) --
std::string Workspace::root_path(const std::string &name) { return home() + "/.semtools/workspaces/" + name; }
std::string Workspace::config_path_for(const std::string &name) { return root_path(name) + "/config.json"; }
std::string Workspace::active(const std::optional<std::string> &workspace_name) {
  std::string a;
  if (workspace_name) a = *workspace_name;
  else if (const char *e = std::getenv("SEMTOOLS_WORKSPACE")) a = e;
  if (a.empty()) throw std::runtime_error("No active workspace. Run: workspace use <name>");
  return a;
}
Workspace Workspace::open(const std::optional<std::string> &workspace_name) {
  const std::string a = active(workspace_name);
  Workspace ws;
  try {
    Json j = Json::parse(read_file(config_path_for(a)));
    const Json *n = j.get("name"), *r = j.get("root_dir"), *b = j.get("in_batch_size"), *o = j.get("oversample_factor");
    if (!n || !r || !b || !o || n->type != Json::Str || r->type != Json::Str) throw std::runtime_error("cfg");
    ws.config.name = n->str; ws.config.root_dir = r->str;
    ws.config.in_batch_size = (size_t)b->num; ws.config.oversample_factor = (size_t)o->num;
  } catch (const std::exception &) {
    ws.config = WorkspaceConfig();
  }
  if (ws.config.root_dir.empty()) ws.config.root_dir = root_path(a);
  if (ws.config.name.empty() || ws.config.name == "default") ws.config.name = a;
  return ws;
}
void Workspace::save() const {
  const std::string p = config_path_for(config.name);
  mkdirs(p.substr(0, p.rfind('/')));
  std::ofstream f(p);
  f << "{\n  \"name\": " << json_string(config.name) << ",\n  \"root_dir\": " << json_string(config.root_dir)
    << ",\n  \"in_batch_size\": " << config.in_batch_size << ",\n  \"oversample_factor\": " << config.oversample_factor << "\n}";
}

// ------------------------------------------------------------------ ids ------------------
uint64_t DocMeta::id() const { return stb_fnv1a64(reinterpret_cast<const uint8_t *>(path.data()), path.size()); }
uint64_t LineEmbedding::id() const { return stb_line_id(reinterpret_cast<const uint8_t *>(path.data()), path.size(), line_number); }

// ------------------------------------------------------------------ Store ----------------
Store Store::open(const std::string &workspace_dir) {
  Store s;
  s.dir_ = workspace_dir + "/flat.b200";
  mkdirs(s.dir_);
  const std::string meta = s.dir_ + "/store.json";
  if (!exists(meta)) return s;
  Json j = Json::parse(read_file(meta));
  const Json *fmt = j.get("format");
  if (!fmt || fmt->str != "semtools_b200.flat.v1") throw std::runtime_error("unknown store format");
  for (const auto &p : j.get("paths")->arr) { s.path_idx_[p.str] = (int32_t)s.paths_.size(); s.paths_.push_back(p.str); }
  for (const auto &d : j.get("docs")->arr) {
    DocMeta m;
    m.path = d.get("path")->str;
    m.size_bytes = std::strtoull(d.get("size_bytes")->raw_num.c_str(), nullptr, 10);
    m.mtime = std::strtoll(d.get("mtime")->raw_num.c_str(), nullptr, 10);
    m.version = (uint32_t)d.get("_version")->num;
    s.docs_.push_back(m);
  }
  const std::string rows = read_file(s.dir_ + "/rows.i32"), emb = read_file(s.dir_ + "/line_embeddings.f32");
  s.rows_.resize(rows.size() / 4);
  std::memcpy(s.rows_.data(), rows.data(), s.rows_.size() * 4);
  s.emb_.resize(emb.size() / 4);
  std::memcpy(s.emb_.data(), emb.data(), s.emb_.size() * 4);
  if (s.rows_.size() / 2 != s.emb_.size() / LINE_EMBEDDING_SIZE) throw std::runtime_error("store files disagree on the row count");
  for (size_t r = 0; r < s.rows_.size() / 2; ++r) {
    LineEmbedding le{s.paths_[s.rows_[2 * r]], s.rows_[2 * r + 1], {}};
    s.id_row_[le.id()] = r;
  }
  s.n_disk_ = s.rows_.size() / 2;
  s.rewrite_ = false;
  return s;
}

void Store::flush() const {
  const std::string tmp = dir_ + "/store.json.tmp";
  {
    std::ofstream f(tmp);
    f << "{\"format\": \"semtools_b200.flat.v1\", \"dim\": 256, \"rows\": " << rows_.size() / 2 << ", \"paths\": [";
    for (size_t i = 0; i < paths_.size(); ++i) f << (i ? ", " : "") << json_string(paths_[i]);
    f << "], \"docs\": [";
    for (size_t i = 0; i < docs_.size(); ++i)
      f << (i ? ", " : "") << "{\"path\": " << json_string(docs_[i].path) << ", \"size_bytes\": " << docs_[i].size_bytes
        << ", \"mtime\": " << docs_[i].mtime << ", \"_version\": " << docs_[i].version << "}";
    f << "]}";
  }
  // the two row files are appended to / patched in place; only deletions rewrite them
  const std::string rows_p = dir_ + "/rows.i32", emb_p = dir_ + "/line_embeddings.f32";
  const size_t n = rows_.size() / 2;
  auto file_size = [](const std::string &p) -> long long { struct stat st; return ::stat(p.c_str(), &st) == 0 ? (long long)st.st_size : -1; };
  if (rewrite_ || file_size(rows_p) != (long long)(n_disk_ * 8) || file_size(emb_p) != (long long)(n_disk_ * LINE_EMBEDDING_SIZE * 4)) {
    { std::ofstream f(rows_p, std::ios::binary); f.write(reinterpret_cast<const char *>(rows_.data()), (std::streamsize)(rows_.size() * 4)); }
    { std::ofstream f(emb_p, std::ios::binary); f.write(reinterpret_cast<const char *>(emb_.data()), (std::streamsize)(emb_.size() * 4)); }
    ++full_rewrites;
  } else {
    std::fstream fr(rows_p, std::ios::in | std::ios::out | std::ios::binary), fe(emb_p, std::ios::in | std::ios::out | std::ios::binary);
    for (size_t r : dirty_) {
      if (r >= n_disk_) continue;
      fr.seekp((std::streamoff)(r * 8)); fr.write(reinterpret_cast<const char *>(rows_.data() + 2 * r), 8);
      fe.seekp((std::streamoff)(r * LINE_EMBEDDING_SIZE * 4));
      fe.write(reinterpret_cast<const char *>(emb_.data() + r * LINE_EMBEDDING_SIZE), LINE_EMBEDDING_SIZE * 4);
    }
    if (n > n_disk_) {
      fr.seekp(0, std::ios::end); fr.write(reinterpret_cast<const char *>(rows_.data() + 2 * n_disk_), (std::streamsize)((n - n_disk_) * 8));
      fe.seekp(0, std::ios::end);
      fe.write(reinterpret_cast<const char *>(emb_.data() + n_disk_ * LINE_EMBEDDING_SIZE), (std::streamsize)((n - n_disk_) * LINE_EMBEDDING_SIZE * 4));
    }
  }
  n_disk_ = n; rewrite_ = false; dirty_.clear();
  std::rename(tmp.c_str(), (dir_ + "/store.json").c_str());
}

std::map<std::string, DocMeta> Store::get_existing_docs(const std::vector<std::string> &paths) const {
  std::map<std::string, DocMeta> out;
  for (const auto &p : paths)
    for (const auto &d : docs_) if (d.path == p) { out[p] = d; break; }
  return out;
}

std::vector<DocumentState> Store::analyze_document_states(const std::vector<std::string> &paths) const {
  auto existing = get_existing_docs(paths);
  std::vector<DocumentState> states;
  for (const auto &fp : paths) {
    struct stat st;
    if (::stat(fp.c_str(), &st) != 0) continue;                     // :578-581 missing file: skipped
    DocMeta cur{fp, (uint64_t)st.st_size, (int64_t)st.st_mtime, CURRENT_EMBEDDING_VERSION};
    DocumentState ds;
    ds.filename = fp; ds.meta = cur;
    auto it = existing.find(fp);
    if (it != existing.end() && it->second.size_bytes == cur.size_bytes && it->second.mtime == cur.mtime &&
        it->second.version == CURRENT_EMBEDDING_VERSION) {
      ds.kind = DocumentState::Unchanged;
    } else {
      ds.kind = it != existing.end() ? DocumentState::Changed : DocumentState::New;
      ds.content = read_file(fp);
    }
    states.push_back(std::move(ds));
  }
  return states;
}

void Store::upsert_document_metadata(const std::vector<DocMeta> &metas) {
  if (metas.empty()) return;
  for (const auto &m : metas) {
    auto it = std::find_if(docs_.begin(), docs_.end(), [&](const DocMeta &d) { return d.path == m.path; });
    if (it != docs_.end()) *it = m; else docs_.push_back(m);
  }
  flush();
}

void Store::upsert_line_embeddings(const std::vector<LineEmbedding> &lines) {
  if (lines.empty()) return;
  for (const auto &le : lines) {
    if (le.embedding.size() != LINE_EMBEDDING_SIZE) throw std::runtime_error("embedding must have 256 floats");
    int32_t pi;
    auto pit = path_idx_.find(le.path);
    if (pit == path_idx_.end()) { pi = (int32_t)paths_.size(); paths_.push_back(le.path); path_idx_[le.path] = pi; }
    else pi = pit->second;
    const uint64_t rid = le.id();
    auto rit = id_row_.find(rid);
    size_t row;
    if (rit != id_row_.end()) { row = rit->second; dirty_.insert(row); }   // upsert replaces by id
    else { row = rows_.size() / 2; id_row_[rid] = row; rows_.resize(rows_.size() + 2); emb_.resize(emb_.size() + LINE_EMBEDDING_SIZE); }
    rows_[2 * row] = pi; rows_[2 * row + 1] = le.line_number;
    std::memcpy(emb_.data() + row * LINE_EMBEDDING_SIZE, le.embedding.data(), LINE_EMBEDDING_SIZE * 4);
  }
  flush();
}

void Store::delete_document_metadata(const std::vector<std::string> &paths) {
  for (const auto &p : paths)
    docs_.erase(std::remove_if(docs_.begin(), docs_.end(), [&](const DocMeta &d) { return d.path == p && d.version == CURRENT_EMBEDDING_VERSION; }), docs_.end());
  if (!paths.empty()) flush();
}

void Store::delete_line_embeddings(const std::vector<std::string> &paths) {
  if (paths.empty()) return;
  std::vector<char> kill(paths_.size(), 0);
  for (const auto &p : paths) { auto it = path_idx_.find(p); if (it != path_idx_.end()) kill[it->second] = 1; }
  size_t w = 0;
  const size_t n = rows_.size() / 2;
  for (size_t r = 0; r < n; ++r) {
    if (kill[rows_[2 * r]]) continue;
    if (w != r) {
      rows_[2 * w] = rows_[2 * r]; rows_[2 * w + 1] = rows_[2 * r + 1];
      std::memmove(emb_.data() + w * LINE_EMBEDDING_SIZE, emb_.data() + r * LINE_EMBEDDING_SIZE, LINE_EMBEDDING_SIZE * 4);
    }
    ++w;
  }
  rows_.resize(2 * w); emb_.resize(w * LINE_EMBEDDING_SIZE);
  if (w != n) rewrite_ = true;                                      // rows moved: the files are rewritten
  id_row_.clear();
  for (size_t r = 0; r < w; ++r) { LineEmbedding le{paths_[rows_[2 * r]], rows_[2 * r + 1], {}}; id_row_[le.id()] = r; }
  flush();
}

void Store::delete_documents(const std::vector<std::string> &paths) {
  if (paths.empty()) return;
  delete_document_metadata(paths);
  delete_line_embeddings(paths);
}

std::vector<std::string> Store::get_all_document_paths() const {
  std::vector<std::string> out;
  for (const auto &d : docs_) out.push_back(d.path);
  return out;
}

std::vector<uint64_t> Store::ranges_for(const std::vector<std::string> &subset_paths) const {
  std::vector<char> sel(paths_.size(), 0);
  for (const auto &p : subset_paths) { auto it = path_idx_.find(p); if (it != path_idx_.end()) sel[it->second] = 1; }
  std::vector<uint64_t> ranges;
  const size_t n = rows_.size() / 2;
  size_t r = 0;
  while (r < n) {
    if (!sel[rows_[2 * r]]) { ++r; continue; }
    size_t e = r;
    while (e < n && sel[rows_[2 * e]]) ++e;
    ranges.push_back(r); ranges.push_back(e);
    r = e;
  }
  return ranges;
}

std::vector<RankedLine> Store::search_line_embeddings(const std::vector<float> &query, const std::vector<std::string> &subset_paths,
                                                      size_t top_k, std::optional<float> max_distance, int device) {
  std::vector<RankedLine> out;
  if (subset_paths.empty() || top_k == 0) return out;               // :489-491
  auto ranges = ranges_for(subset_paths);
  if (ranges.empty() || query.size() != LINE_EMBEDDING_SIZE) return out;
  stb_ctx *ctx = nullptr; stb_corpus *corpus = nullptr;
  auto chk = [&](int rc) { if (rc < 0) { std::string m = stb_last_error(); stb_corpus_destroy(corpus); stb_ctx_destroy(ctx); throw StbError(rc, m); } };
  chk(stb_ctx_create(device, nullptr, &ctx));
  chk(stb_corpus_create(ctx, STB_DIM, rows_.size() / 2, 0, &corpus));
  chk(stb_corpus_append(corpus, emb_.data(), rows_.size() / 2));
  std::vector<stb_hit> hits(top_k);
  uint64_t n = 0;
  chk(stb_search(ctx, corpus, query.data(), (uint32_t)top_k, max_distance ? 1 : 0, max_distance ? (double)*max_distance : 0.0,
                 STB_MODE_STORE_QUERY, ranges.data(), (uint32_t)(ranges.size() / 2), hits.data(), top_k, &n));
  for (uint64_t i = 0; i < n; ++i)
    out.push_back({paths_[rows_[2 * hits[i].row]], rows_[2 * hits[i].row + 1], (float)hits[i].distance});   // :531 f32
  stb_corpus_destroy(corpus);
  stb_ctx_destroy(ctx);
  return out;
}

// ------------------------------------------------------------------ search_with_workspace --
std::vector<RankedLine> search_with_workspace(const std::vector<std::string> &files, const std::vector<float> &query_embedding,
                                              const EmbedLinesFn &embed_lines, const SearchConfig &cfg,
                                              const std::optional<std::string> &workspace_name,
                                              const std::function<void(const std::string &)> &log, int device) {
  Workspace ws = Workspace::open(workspace_name);
  Store store = Store::open(ws.config.root_dir);
  std::vector<LineEmbedding> to_upsert;
  std::vector<DocMeta> docs;
  for (auto &st : store.analyze_document_states(files)) {
    if (st.kind == DocumentState::Unchanged) continue;
    const std::vector<std::string> lines = rust_lines(st.content);
    if (lines.empty()) continue;                                   // create_document_from_content -> None
    const std::vector<float> emb = embed_lines(lines);
    if (emb.size() != lines.size() * LINE_EMBEDDING_SIZE) throw std::runtime_error("embed_lines returned the wrong size");
    for (size_t i = 0; i < lines.size(); ++i)
      to_upsert.push_back({st.filename, (int32_t)i,                 // 0-based line numbers (mod.rs:178)
                           std::vector<float>(emb.begin() + i * LINE_EMBEDDING_SIZE, emb.begin() + (i + 1) * LINE_EMBEDDING_SIZE)});
    docs.push_back(st.meta);
  }
  if (!to_upsert.empty()) {
    if (log) log("Updating workspace with " + std::to_string(to_upsert.size()) + " lines from new/changed docs...");
    store.upsert_line_embeddings(to_upsert);
  }
  if (!docs.empty()) {
    if (log) log("Updating workspace with " + std::to_string(docs.size()) + " new/changed documents...");
    store.upsert_document_metadata(docs);
  }
  std::optional<float> max_d;
  if (cfg.max_distance) max_d = (float)*cfg.max_distance;           // mod.rs:211 `as f32`
  return store.search_line_embeddings(query_embedding, files, cfg.top_k, max_d, device);
}

static bool read_lines(const std::string &path, std::vector<std::string> &lines) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss; ss << f.rdbuf();
  lines = rust_lines(ss.str());
  return true;
}

std::string format_workspace_search_results(const std::vector<RankedLine> &ranked, size_t n_lines, bool is_tty) {
  std::string out;
  char num[32];
  for (const auto &rl : ranked) {
    const size_t m = (size_t)rl.line_number;
    const size_t start = m > n_lines ? m - n_lines : 0;
    const size_t end = m + n_lines + 1;                             // NOT clamped in the header (:77-79)
    out += rl.path + ":" + std::to_string(start) + "::" + std::to_string(end) + " (" + rust_display_f32(rl.distance) + ")\n";
    std::vector<std::string> lines;
    if (read_lines(rl.path, lines)) {
      const size_t actual_end = std::min(end, lines.size());
      if (start > actual_end) throw std::runtime_error("slice index starts past the end (the reference panics here)");
      for (size_t n = start; n < actual_end; ++n) {
        snprintf(num, sizeof(num), "%4zu", n + 1);
        if (n == m && is_tty) out += std::string("\x1b[43m\x1b[30m") + num + ": " + lines[n] + "\x1b[0m\n";
        else out += std::string(num) + ": " + lines[n] + "\n";
      }
    } else {
      out += "    [Error: Could not read file content]\n";
    }
    out += "\n";
  }
  return out;
}

std::string workspace_output_json(const std::vector<RankedLine> &ranked, size_t n_lines) {
  if (ranked.empty()) return "{\n  \"results\": []\n}";
  std::string o = "{\n  \"results\": [\n";
  for (size_t i = 0; i < ranked.size(); ++i) {
    const auto &rl = ranked[i];
    const size_t m = (size_t)rl.line_number;
    const size_t start = m > n_lines ? m - n_lines : 0, end = m + n_lines + 1;
    std::string content;
    std::vector<std::string> lines;
    if (read_lines(rl.path, lines)) {
      for (size_t n = start; n < std::min(end, lines.size()); ++n) { if (n > start) content += "\n"; content += lines[n]; }
    } else {
      content = "[Error: Could not read file content]";
    }
    o += "    {\n";
    o += "      \"filename\": " + json_string(rl.path) + ",\n";
    o += "      \"start_line_number\": " + std::to_string(start) + ",\n";
    o += "      \"end_line_number\": " + std::to_string(end) + ",\n";
    o += "      \"match_line_number\": " + std::to_string(m) + ",\n";
    o += "      \"distance\": " + json_f64((double)rl.distance) + ",\n";   // f32 widened to f64 (:233)
    o += "      \"content\": " + json_string(content) + "\n";
    o += i + 1 < ranked.size() ? "    },\n" : "    }\n";
  }
  o += "  ]\n}";
  return o;
}

}  // namespace semtools
