// ORACLE (test infrastructure only -- never linked into the product library).
// Minimal reader for the Fortran NAMELIST subset used by SEM2DPACK's Par.inp.
// Mirrors how the reference consumes the file: each reader either rewinds or scans
// FORWARD from the current record for the next "&NAME" group
// (/root/reference/SRC/input.f90:12-63, mat_gen.f90:119-189, bc_gen.f90:98-124), and
// list-directed reads (read(iin,*)) take whole records after the group's closing '/'
// (distribution_pwconr.f90:58-66, distribution_order0.f90:55-69).
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

inline std::string upper(std::string s) {
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}
inline std::string lower(std::string s) {
  for (auto& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}
inline double fortran_to_double(std::string s) {
  for (auto& c : s)
    if (c == 'd' || c == 'D') c = 'e';
  return std::strtod(s.c_str(), nullptr);
}
// Fortran REAL (single) namelist variable widened to double (stf_ricker.f90:53,71-73)
inline double fortran_to_real_as_double(const std::string& s) {
  return (double)(float)fortran_to_double(s);
}
inline bool fortran_to_logical(std::string s) {
  s = upper(s);
  size_t p = (!s.empty() && s[0] == '.') ? 1 : 0;
  return p < s.size() && s[p] == 'T';
}

struct NmlGroup {
  std::string name;                                    // upper case
  std::map<std::string, std::vector<std::string>> kv;  // key (lower case) -> raw values
  int start_line = 0, end_line = 0;
  bool has(const std::string& k) const { return kv.count(lower(k)) != 0; }
  const std::vector<std::string>& raw(const std::string& k) const { return kv.at(lower(k)); }
  double dbl(const std::string& k, double dflt, int idx = 0) const {
    if (!has(k) || (int)raw(k).size() <= idx) return dflt;
    return fortran_to_double(raw(k)[idx]);
  }
  double real_as_dbl(const std::string& k, double dflt) const {
    if (!has(k)) return (double)(float)dflt;
    return fortran_to_real_as_double(raw(k)[0]);
  }
  int integer(const std::string& k, int dflt, int idx = 0) const {
    if (!has(k) || (int)raw(k).size() <= idx) return dflt;
    return (int)std::strtol(raw(k)[idx].c_str(), nullptr, 10);
  }
  std::string str(const std::string& k, const std::string& dflt, int idx = 0) const {
    if (!has(k) || (int)raw(k).size() <= idx) return dflt;
    return raw(k)[idx];
  }
  bool logical(const std::string& k, bool dflt) const {
    if (!has(k)) return dflt;
    return fortran_to_logical(raw(k)[0]);
  }
};

class ParInp {
 public:
  std::vector<std::string> lines;
  std::vector<NmlGroup> groups;  // in file order
  int cursor_line = 0;           // next record to be read (0-based)

  ParInp() {}
  explicit ParInp(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("ParInp: cannot open " + path);
    std::string l;
    while (std::getline(f, l)) lines.push_back(l);
    parse();
  }
  static ParInp from_string(const std::string& text) {
    ParInp p;
    std::istringstream f(text);
    std::string l;
    while (std::getline(f, l)) p.lines.push_back(l);
    p.parse();
    return p;
  }

  void rewind() { cursor_line = 0; }

  // Fortran: read(iin,NAME,END=...) -- next group called NAME at/after the current record.
  const NmlGroup* next(const std::string& name_) {
    std::string name = upper(name_);
    for (const auto& g : groups) {
      if (g.start_line >= cursor_line && g.name == name) {
        cursor_line = g.end_line + 1;
        return &g;
      }
    }
    cursor_line = (int)lines.size();  // END= : positioned at end of file
    return nullptr;
  }
  int count(const std::string& name_) const {
    std::string name = upper(name_);
    int n = 0;
    for (const auto& g : groups) n += (g.name == name);
    return n;
  }

  // Fortran: read(iin,*) v(1:n) -- list-directed, consumes whole records
  std::vector<double> read_list(int n) {
    std::vector<double> out;
    while ((int)out.size() < n) {
      if (cursor_line >= (int)lines.size()) throw std::runtime_error("ParInp: EOF in list read");
      std::string l = lines[cursor_line++];
      for (auto& c : l)
        if (c == ',') c = ' ';
      std::istringstream ss(l);
      std::string tok;
      while ((int)out.size() < n && ss >> tok) {
        if (tok[0] == '#' || tok[0] == '!' || tok[0] == '/') break;
        out.push_back(fortran_to_double(tok));
      }
    }
    return out;
  }

 private:
  static bool is_key_start(const std::string& s, size_t p) {
    // identifier followed (after blanks) by '='
    if (!(std::isalpha((unsigned char)s[p]) || s[p] == '_')) return false;
    size_t q = p;
    while (q < s.size() && (std::isalnum((unsigned char)s[q]) || s[q] == '_')) ++q;
    while (q < s.size() && std::isspace((unsigned char)s[q])) ++q;
    return q < s.size() && s[q] == '=';
  }

  void parse() {
    int nl = (int)lines.size();
    int i = 0;
    while (i < nl) {
      const std::string& l = lines[i];
      size_t p = l.find_first_not_of(" \t");
      if (p == std::string::npos || l[p] != '&') {
        ++i;
        continue;
      }
      // group starts here: collect text up to the closing '/' outside quotes
      NmlGroup g;
      g.start_line = i;
      std::string body;
      bool closed = false;
      char quote = 0;
      int li = i;
      size_t pos = p + 1;
      while (li < nl && !closed) {
        const std::string& s = lines[li];
        for (; pos < s.size(); ++pos) {
          char c = s[pos];
          if (quote) {
            if (c == quote) quote = 0;
            body.push_back(c);
          } else if (c == '\'' || c == '"') {
            quote = c;
            body.push_back(c);
          } else if (c == '/') {
            closed = true;
            break;
          } else if (c == '!') {
            break;  // namelist comment
          } else {
            body.push_back(c);
          }
        }
        if (!closed) {
          body.push_back(' ');
          ++li;
          pos = 0;
        }
      }
      g.end_line = std::min(li, nl - 1);
      // name
      size_t q = 0;
      while (q < body.size() && !std::isspace((unsigned char)body[q])) ++q;
      g.name = upper(body.substr(0, q));
      // key = values
      while (q < body.size()) {
        while (q < body.size() && (std::isspace((unsigned char)body[q]) || body[q] == ',')) ++q;
        if (q >= body.size()) break;
        if (!is_key_start(body, q)) {  // stray token: skip
          ++q;
          continue;
        }
        size_t k0 = q;
        while (q < body.size() && (std::isalnum((unsigned char)body[q]) || body[q] == '_')) ++q;
        std::string key = lower(body.substr(k0, q - k0));
        while (body[q] != '=') ++q;
        ++q;
        std::vector<std::string> vals;
        while (q < body.size()) {
          while (q < body.size() && (std::isspace((unsigned char)body[q]) || body[q] == ',')) ++q;
          if (q >= body.size()) break;
          if (body[q] == '\'' || body[q] == '"') {
            char qc = body[q];
            size_t e = body.find(qc, q + 1);
            if (e == std::string::npos) e = body.size();
            vals.push_back(body.substr(q + 1, e - q - 1));
            q = e + 1;
          } else {
            if (is_key_start(body, q)) {
              // T / F logicals are also identifiers; a key needs '=' which is_key_start checked
              break;
            }
            size_t e = q;
            while (e < body.size() && !std::isspace((unsigned char)body[e]) && body[e] != ',') ++e;
            vals.push_back(body.substr(q, e - q));
            q = e;
          }
        }
        g.kv[key] = vals;
      }
      groups.push_back(g);
      i = g.end_line + 1;
    }
  }
};

}  // namespace orc
