// Reader for the Fortran NAMELIST subset of SEM2DPACK's Par.inp, for the C++ host of the B200 path.
// The reference reads the file with `read(iin, GROUP, END=..)`, which scans FORWARD from the current
// record for the next "&GROUP" (SRC/input.f90:12-63; sub-blocks such as &MAT_ELASTIC or &BC_ABSORB
// are read right after their parent, SRC/mat_gen.f90:119-189, SRC/bc_gen.f90:98-124), or rewinds
// first (SRC/time.f90:143, SRC/receivers.f90:85).  `find(name, from)` / `rewind()` give the same two
// access patterns.  Comment lines start with '#' or '!'.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sem2d {

struct nml_group {
  std::string name;                                         // upper case, without '&'
  std::map<std::string, std::vector<std::string>> items;    // lower-case key -> value tokens
  size_t end_pos = 0;                                       // offset of the closing '/' in namelist_file::text()
  bool has(const std::string& k) const { return items.count(k) != 0; }
  static double to_double(std::string t) {
    for (char& c : t)
      if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(t.c_str(), nullptr);
  }
  double real8(const std::string& k, double dflt, size_t idx = 0) const {
    auto it = items.find(k);
    return (it == items.end() || it->second.size() <= idx) ? dflt : to_double(it->second[idx]);
  }
  // a default-kind REAL namelist variable widened with dble() (e.g. SRC/stf_ricker.f90:53,71-73)
  double real4(const std::string& k, double dflt) const { return (double)(float)real8(k, dflt); }
  int integer(const std::string& k, int dflt, size_t idx = 0) const {
    auto it = items.find(k);
    return (it == items.end() || it->second.size() <= idx) ? dflt : (int)std::strtol(it->second[idx].c_str(), nullptr, 10);
  }
  bool logical(const std::string& k, bool dflt) const {
    auto it = items.find(k);
    if (it == items.end() || it->second.empty()) return dflt;
    for (char c : it->second[0])
      if (c != '.') return c == 'T' || c == 't';
    return dflt;
  }
  std::string text(const std::string& k, const std::string& dflt, size_t idx = 0) const {
    auto it = items.find(k);
    return (it == items.end() || it->second.size() <= idx) ? dflt : it->second[idx];
  }
  size_t count(const std::string& k) const {
    auto it = items.find(k);
    return it == items.end() ? 0 : it->second.size();
  }
};

class namelist_file {
 public:
  explicit namelist_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    std::string line;
    while (std::getline(f, line)) {
      size_t p = line.find_first_not_of(" \t");
      if (p != std::string::npos && (line[p] == '#' || line[p] == '!')) continue;
      ss << line << '\n';
    }
    text_ = ss.str();
    parse(text_);
  }
  size_t size() const { return groups_.size(); }
  const std::string& text() const { return text_; }
  // The list-directed records (read(iin,*)) that follow group k: the next n numbers, starting on the line
  // after the one that closes the group (e.g. SRC/distribution_order0.f90:55-69, distribution_pwconr.f90:36-40)
  std::vector<double> records_after(size_t k, size_t n) const {
    size_t p = text_.find('\n', groups_[k].end_pos);
    std::vector<double> v;
    if (p == std::string::npos) return v;
    std::stringstream ss(text_.substr(p + 1));
    std::string tok;
    while (v.size() < n && ss >> tok) {
      if (tok[0] == '&') break;
      v.push_back(nml_group::to_double(tok));
    }
    return v;
  }
  const nml_group& at(size_t k) const { return groups_[k]; }
  // index of the first group called `name` at or after `from`, or -1 (END= branch of the read)
  long find(const std::string& name, size_t from = 0) const {
    for (size_t k = from; k < groups_.size(); ++k)
      if (groups_[k].name == name) return (long)k;
    return -1;
  }

 private:
  std::vector<nml_group> groups_;
  std::string text_;
  static bool key_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }
  void parse(const std::string& s) {
    size_t p = 0;
    const size_t n = s.size();
    while (p < n) {
      if (s[p] != '&') {
        ++p;
        continue;
      }
      nml_group g;
      ++p;
      while (p < n && key_char(s[p])) g.name += (char)std::toupper((unsigned char)s[p++]);
      std::string key;
      while (p < n && s[p] != '/') {
        const char c = s[p];
        if (std::isspace((unsigned char)c) || c == ',') {
          ++p;
        } else if (c == '\'' || c == '"') {  // quoted string value
          std::string v;
          for (++p; p < n && s[p] != c; ++p) v += s[p];
          ++p;
          if (!key.empty()) g.items[key].push_back(v);
        } else {  // bare token: either "key =" or a value
          std::string t;
          while (p < n && !std::isspace((unsigned char)s[p]) && s[p] != ',' && s[p] != '=' && s[p] != '/') t += s[p++];
          size_t q = p;
          while (q < n && (s[q] == ' ' || s[q] == '\t')) ++q;
          if (q < n && s[q] == '=') {
            key.clear();
            for (char ch : t) key += (char)std::tolower((unsigned char)ch);
            g.items[key];
            p = q + 1;
          } else if (!key.empty()) {
            g.items[key].push_back(t);
          }
        }
      }
      g.end_pos = p;
      ++p;  // '/'
      groups_.push_back(g);
    }
  }
};

}  // namespace sem2d
