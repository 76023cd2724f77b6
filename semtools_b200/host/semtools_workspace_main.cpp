// `semtools_b200_workspace`: C++ mirror of the reference's `workspace use|status|prune`
// commands (src/cmds/workspace.rs:11-176) over the flat store, plus two store utilities the
// interop tests use:
//   store-dump <workspace_dir>        canonical text dump of a store (docs, rows, row checksums)
//   store-selftest <workspace_dir>    builds a store through every mutating call and dumps it
// None of these touches the GPU (the store query is Store::search_line_embeddings, used by the
// search CLI).
#include <sys/stat.h>

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>

#include "semtools_store.hpp"

using namespace semtools;

static void dump(const Store &s) {
  auto paths = s.get_all_document_paths();
  auto docs = s.get_existing_docs(paths);
  printf("documents %zu\n", s.count_documents());
  for (const auto &p : paths) {
    const DocMeta &m = docs.at(p);
    printf("doc %s size=%" PRIu64 " mtime=%" PRId64 " version=%u id=%" PRIu64 "\n", m.path.c_str(), m.size_bytes, m.mtime, m.version, m.id());
  }
  printf("line_embeddings %zu\n", s.count_line_embeddings());
  const auto &emb = s.matrix();
  auto all = s.ranges_for(paths);                    // rows of documents that still have metadata
  printf("ranges");
  for (auto r : all) printf(" %" PRIu64, r);
  printf("\n");
  for (size_t r = 0; r < s.count_line_embeddings(); ++r) {
    uint64_t h = stb_fnv1a64(reinterpret_cast<const uint8_t *>(emb.data() + r * LINE_EMBEDDING_SIZE), LINE_EMBEDDING_SIZE * 4);
    printf("row %zu fnv=%" PRIu64 "\n", r, h);
  }
}

static std::vector<float> fake_embedding(const std::string &path, int32_t line, float scale) {
  std::vector<float> e(LINE_EMBEDDING_SIZE);
  uint64_t h = stb_line_id(reinterpret_cast<const uint8_t *>(path.data()), path.size(), line);
  for (size_t i = 0; i < LINE_EMBEDDING_SIZE; ++i) {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    e[i] = scale * (float)((int64_t)(h >> 40) - (1 << 23)) / (float)(1 << 23);
  }
  return e;
}

static int store_selftest(const std::string &dir) {
  Store s = Store::open(dir);
  s.upsert_document_metadata({{"a.txt", 10, 1700000000, CURRENT_EMBEDDING_VERSION}, {"dir/b \"q\".txt", 20, 1700000001, CURRENT_EMBEDDING_VERSION},
                              {"c\xC3\xA9.txt", 30, 1700000002, CURRENT_EMBEDDING_VERSION}, {"old.txt", 5, 1600000000, 1}});
  std::vector<LineEmbedding> ls;
  for (int i = 0; i < 3; ++i) ls.push_back({"a.txt", i, fake_embedding("a.txt", i, 1.f)});
  for (int i = 0; i < 2; ++i) ls.push_back({"dir/b \"q\".txt", i, fake_embedding("dir/b \"q\".txt", i, 1.f)});
  for (int i = 0; i < 4; ++i) ls.push_back({"c\xC3\xA9.txt", i, fake_embedding("c\xC3\xA9.txt", i, 1.f)});
  ls.push_back({"old.txt", 0, fake_embedding("old.txt", 0, 1.f)});
  s.upsert_line_embeddings(ls);
  s.upsert_line_embeddings({{"a.txt", 1, fake_embedding("a.txt", 1, 0.5f)}});          // replace by id
  s.upsert_document_metadata({{"a.txt", 11, 1700000005, CURRENT_EMBEDDING_VERSION}});   // replace in place
  s.delete_documents({"dir/b \"q\".txt", "old.txt"});          // old.txt: version 1 metadata stays, lines go
  s.upsert_line_embeddings({{"a.txt", 3, fake_embedding("a.txt", 3, 1.f)}});            // appended after c's rows
  Store again = Store::open(dir);                              // what is on disk
  dump(again);
  return 0;
}

static int real_main(int argc, char **argv) {
  try {
    std::vector<std::string> a(argv + 1, argv + argc);
    bool json = false;
    std::optional<std::string> wsname;
    std::vector<std::string> pos;
    for (size_t i = 0; i < a.size(); ++i) {
      if (a[i] == "--json" || a[i] == "-j") json = true;
      else if (a[i] == "--workspace" && i + 1 < a.size()) wsname = a[++i];
      else pos.push_back(a[i]);
    }
    if (pos.size() == 2 && pos[0] == "store-dump") { dump(Store::open(pos[1])); return 0; }
    if (pos.size() == 2 && pos[0] == "store-selftest") return store_selftest(pos[1]);
    if (pos.size() == 2 && pos[0] == "format-ranked") {
      // test hook: stdin lines `path<TAB>line_number<TAB>f32 bits (decimal u32)` -> the workspace renderers
      std::vector<RankedLine> ranked;
      char buf[8192];
      while (fgets(buf, sizeof(buf), stdin)) {
        std::string l(buf);
        if (!l.empty() && l.back() == '\n') l.pop_back();
        const size_t t1 = l.find('\t'), t2 = l.find('\t', t1 + 1);
        if (t1 == std::string::npos || t2 == std::string::npos) continue;
        const uint32_t bits = (uint32_t)std::stoul(l.substr(t2 + 1));
        float d;
        memcpy(&d, &bits, 4);
        ranked.push_back({l.substr(0, t1), (int32_t)std::stol(l.substr(t1 + 1, t2 - t1 - 1)), d});
      }
      const size_t n_lines = std::stoul(pos[1]);
      if (json) printf("%s\n", workspace_output_json(ranked, n_lines).c_str());
      else fputs(format_workspace_search_results(ranked, n_lines, false).c_str(), stdout);
      return 0;
    }
    if (pos.size() == 2 && pos[0] == "use") {                          // workspace.rs:11-67
      Workspace ws;
      ws.config.name = pos[1];
      ws.config.root_dir = Workspace::root_path(pos[1]);
      ws.save();
      if (json) {
        size_t total = 0;
        try { total = Store::open(ws.config.root_dir).get_stats().total_documents; } catch (const std::exception &) {}
        printf("{\n  \"name\": %s,\n  \"root_dir\": %s,\n  \"total_documents\": %zu\n}\n", json_string(ws.config.name).c_str(),
               json_string(ws.config.root_dir).c_str(), total);
      } else {
        printf("Workspace '%s' configured.\nTo activate it, run:\n  export SEMTOOLS_WORKSPACE=%s\n\n"
               "Or add this to your shell profile (.bashrc, .zshrc, etc.)\n\n"
               "Or use the `--workspace` option on the commands that support it\n", pos[1].c_str(), pos[1].c_str());
      }
      return 0;
    }
    if (pos.size() == 1 && pos[0] == "status") {                       // workspace.rs:69-113
      Workspace::active(wsname);
      Workspace ws = Workspace::open(wsname);
      WorkspaceStats st = Store::open(ws.config.root_dir).get_stats();
      if (json) {
        printf("{\n  \"name\": %s,\n  \"root_dir\": %s,\n  \"total_documents\": %zu\n}\n", json_string(ws.config.name).c_str(),
               json_string(ws.config.root_dir).c_str(), st.total_documents);
      } else {
        printf("Active workspace: %s\nRoot: %s\nDocuments: %zu\n", ws.config.name.c_str(), ws.config.root_dir.c_str(), st.total_documents);
        if (st.has_index) printf("Index: Yes (%s)\n", st.index_type.empty() ? "Unknown" : st.index_type.c_str());
        else printf("Index: No\n");
      }
      return 0;
    }
    if (pos.size() == 1 && pos[0] == "prune") {                        // workspace.rs:115-176
      Workspace::active(wsname);
      Workspace ws = Workspace::open(wsname);
      Store store = Store::open(ws.config.root_dir);
      auto all = store.get_all_document_paths();
      std::vector<std::string> missing;
      for (const auto &p : all) { struct stat sb; if (::stat(p.c_str(), &sb) != 0) missing.push_back(p); }
      if (!missing.empty()) store.delete_documents(missing);
      if (json) {
        printf("{\n  \"files_removed\": %zu,\n  \"files_remaining\": %zu\n}\n", missing.size(), all.size() - missing.size());
      } else if (missing.empty()) {
        printf("No stale documents found. Workspace is clean.\n");
      } else {
        printf("Found %zu stale documents:\n", missing.size());
        for (const auto &p : missing) printf("  - %s\n", p.c_str());
        printf("Removed %zu stale documents from workspace.\n", missing.size());
      }
      return 0;
    }
    fprintf(stderr, "usage: semtools_b200_workspace [--json] [--workspace NAME] use <name> | status | prune | store-dump <dir> | store-selftest <dir>\n");
    return 2;
  } catch (const std::exception &e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
}

// nothing a bad input file can throw ends in std::terminate: every exit is a message + status
int main(int argc, char **argv) {
  try {
    return real_main(argc, argv);
  } catch (const std::exception &e) {
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  } catch (...) {
    fprintf(stderr, "Error: unknown failure\n");
    return 1;
  }
}
