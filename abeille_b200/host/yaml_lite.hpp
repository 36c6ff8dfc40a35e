/* yaml_lite.hpp -- the subset of YAML that Abeille input decks use.
 *
 * The reference reads its decks with yaml-cpp (src/parser.cpp:77-137), which is not available
 * here; this reader covers what the shipped decks (and PyYAML re-dumps of them) contain:
 * block mappings and sequences, flow sequences / mappings (possibly spanning lines), comments,
 * plain, single- and double-quoted scalars.  Anchors, tags, multi-documents and block scalars
 * are not supported and raise.  Numbers are converted with strtod, which rounds correctly like
 * yaml-cpp's stream extraction does.
 */
#pragma once
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace yaml_lite {

class Node {
 public:
  enum Type { Null, Scalar, Sequence, Map };
  Type type = Null;
  std::string scalar;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node>> map;

  bool IsDefined() const { return type != Null; }
  explicit operator bool() const { return type != Null; }
  bool IsSequence() const { return type == Sequence; }
  bool IsMap() const { return type == Map; }
  bool IsScalar() const { return type == Scalar; }
  size_t size() const { return type == Sequence ? seq.size() : (type == Map ? map.size() : 0); }

  const Node& operator[](const std::string& key) const {
    static const Node null_node;
    if (type != Map) return null_node;
    for (const auto& kv : map)
      if (kv.first == key) return kv.second;
    return null_node;
  }
  const Node& operator[](const char* key) const { return (*this)[std::string(key)]; }
  const Node& operator[](size_t i) const {
    static const Node null_node;
    if (type != Sequence || i >= seq.size()) return null_node;
    return seq[i];
  }
  const Node& operator[](int i) const { return (*this)[static_cast<size_t>(i)]; }

  double as_double() const {
    if (type != Scalar) throw std::runtime_error("yaml: expected a number");
    const char* s = scalar.c_str();
    char* end = nullptr;
    const double v = std::strtod(s, &end);
    if (end == s) throw std::runtime_error("yaml: '" + scalar + "' is not a number");
    return v;
  }
  long long as_int() const {
    if (type != Scalar) throw std::runtime_error("yaml: expected an integer");
    const char* s = scalar.c_str();
    char* end = nullptr;
    const long long v = std::strtoll(s, &end, 10);
    if (end == s) throw std::runtime_error("yaml: '" + scalar + "' is not an integer");
    if (*end == '.' || *end == 'e' || *end == 'E') return static_cast<long long>(as_double());
    return v;
  }
  unsigned long long as_uint64() const {
    if (type != Scalar) throw std::runtime_error("yaml: expected an integer");
    return std::strtoull(scalar.c_str(), nullptr, 10);
  }
  bool as_bool() const {
    if (type != Scalar) throw std::runtime_error("yaml: expected a boolean");
    const std::string& s = scalar;
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on") return true;
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off") return false;
    throw std::runtime_error("yaml: '" + s + "' is not a boolean");
  }
  const std::string& as_string() const {
    if (type != Scalar) throw std::runtime_error("yaml: expected a string");
    return scalar;
  }
  std::vector<double> as_doubles() const {
    if (type != Sequence) throw std::runtime_error("yaml: expected a sequence of numbers");
    std::vector<double> v;
    v.reserve(seq.size());
    for (const auto& n : seq) v.push_back(n.as_double());
    return v;
  }
};

namespace detail {

struct Line {
  int indent;
  std::string text;  // without indentation, comments stripped, flow collections joined
  int number;
};

inline std::string rtrim(std::string s) {
  while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r' || s.back() == '\n')) s.pop_back();
  return s;
}
inline std::string ltrim(const std::string& s) {
  size_t i = 0;
  while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
  return s.substr(i);
}

// strips a trailing comment and reports the bracket depth change of the line
inline std::string strip_comment(const std::string& in, int& depth) {
  std::string out;
  char quote = 0;
  for (size_t i = 0; i < in.size(); i++) {
    const char c = in[i];
    if (quote) {
      out += c;
      if (c == quote) quote = 0;
      continue;
    }
    if (c == '"' || c == '\'') {
      quote = c;
      out += c;
      continue;
    }
    if (c == '#' && (i == 0 || in[i - 1] == ' ' || in[i - 1] == '\t')) break;
    if (c == '[' || c == '{') depth++;
    if (c == ']' || c == '}') depth--;
    out += c;
  }
  return rtrim(out);
}

inline std::vector<Line> logical_lines(std::istream& in) {
  std::vector<Line> lines;
  std::string raw;
  int number = 0, depth = 0;
  while (std::getline(in, raw)) {
    number++;
    if (depth > 0) {  // continuation of a flow collection
      const std::string t = ltrim(strip_comment(raw, depth));
      if (!t.empty()) lines.back().text += " " + t;
      continue;
    }
    int d = 0;
    std::string t = strip_comment(raw, d);
    if (ltrim(t).empty()) continue;
    if (ltrim(t) == "---") continue;
    int indent = 0;
    while (indent < static_cast<int>(t.size()) && t[static_cast<size_t>(indent)] == ' ') indent++;
    if (indent < static_cast<int>(t.size()) && t[static_cast<size_t>(indent)] == '\t')
      throw std::runtime_error("yaml: tab indentation at line " + std::to_string(number));
    lines.push_back({indent, t.substr(static_cast<size_t>(indent)), number});
    depth = d;
  }
  if (depth != 0) throw std::runtime_error("yaml: unbalanced brackets at end of file");
  return lines;
}

class FlowParser {
 public:
  explicit FlowParser(const std::string& s) : s_(s) {}
  Node parse() {
    Node n = value(false);
    skip();
    if (i_ != s_.size()) throw std::runtime_error("yaml: trailing characters in '" + s_ + "'");
    return n;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  void skip() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\t')) i_++;
  }
  Node scalar(bool in_flow, bool is_key) {
    Node n;
    n.type = Node::Scalar;
    skip();
    if (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) {
      const char q = s_[i_++];
      while (i_ < s_.size() && s_[i_] != q) {
        if (q == '"' && s_[i_] == '\\' && i_ + 1 < s_.size()) i_++;
        n.scalar += s_[i_++];
      }
      if (i_ >= s_.size()) throw std::runtime_error("yaml: unterminated string in '" + s_ + "'");
      i_++;
      return n;
    }
    const size_t start = i_;
    while (i_ < s_.size()) {
      const char c = s_[i_];
      if (in_flow && (c == ',' || c == ']' || c == '}')) break;
      if (is_key && c == ':' && (i_ + 1 == s_.size() || s_[i_ + 1] == ' ')) break;
      i_++;
    }
    n.scalar = rtrim(s_.substr(start, i_ - start));
    if (n.scalar == "~" || n.scalar == "null" || n.scalar.empty()) n.type = Node::Null;
    return n;
  }
  Node value(bool in_flow) {
    skip();
    if (i_ >= s_.size()) return Node();
    if (s_[i_] == '[') {
      i_++;
      Node n;
      n.type = Node::Sequence;
      for (;;) {
        skip();
        if (i_ >= s_.size()) throw std::runtime_error("yaml: unterminated '[' in '" + s_ + "'");
        if (s_[i_] == ']') { i_++; break; }
        n.seq.push_back(value(true));
        skip();
        if (i_ < s_.size() && s_[i_] == ',') i_++;
      }
      return n;
    }
    if (s_[i_] == '{') {
      i_++;
      Node n;
      n.type = Node::Map;
      for (;;) {
        skip();
        if (i_ >= s_.size()) throw std::runtime_error("yaml: unterminated '{' in '" + s_ + "'");
        if (s_[i_] == '}') { i_++; break; }
        Node key = scalar(true, true);
        skip();
        if (i_ >= s_.size() || s_[i_] != ':') throw std::runtime_error("yaml: expected ':' in '" + s_ + "'");
        i_++;
        Node val = value(true);
        n.map.emplace_back(key.scalar, std::move(val));
        skip();
        if (i_ < s_.size() && s_[i_] == ',') i_++;
      }
      return n;
    }
    return scalar(in_flow, false);
  }
};

// position of the "key:" separator of a block mapping entry, or npos
inline size_t find_key_colon(const std::string& t) {
  if (t.empty() || t[0] == '[' || t[0] == '{') return std::string::npos;
  char quote = 0;
  for (size_t i = 0; i < t.size(); i++) {
    const char c = t[i];
    if (quote) {
      if (c == quote) quote = 0;
      continue;
    }
    if (c == '"' || c == '\'') {
      quote = c;
      continue;
    }
    if (c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
  }
  return std::string::npos;
}

class BlockParser {
 public:
  explicit BlockParser(std::vector<Line> lines) : L_(std::move(lines)) {}
  Node parse() {
    if (L_.empty()) return Node();
    Node n = block(L_[0].indent);
    if (pos_ != L_.size()) throw std::runtime_error("yaml: bad indentation at line " + std::to_string(L_[pos_].number));
    return n;
  }

 private:
  std::vector<Line> L_;
  size_t pos_ = 0;

  static bool is_seq_item(const std::string& t) { return t[0] == '-' && (t.size() == 1 || t[1] == ' '); }

  Node inline_value(const std::string& text) { return FlowParser(text).parse(); }

  Node block(int indent) {
    if (pos_ >= L_.size()) return Node();
    if (is_seq_item(L_[pos_].text)) return sequence(indent);
    if (find_key_colon(L_[pos_].text) != std::string::npos) return mapping(indent);
    Node n = inline_value(L_[pos_].text);
    pos_++;
    return n;
  }

  Node sequence(int indent) {
    Node n;
    n.type = Node::Sequence;
    while (pos_ < L_.size() && L_[pos_].indent == indent && is_seq_item(L_[pos_].text)) {
      const std::string rest = L_[pos_].text.substr(1);
      const std::string content = ltrim(rest);
      if (content.empty()) {  // "-" alone: the item is the nested block below
        pos_++;
        if (pos_ < L_.size() && L_[pos_].indent > indent) n.seq.push_back(block(L_[pos_].indent));
        else n.seq.push_back(Node());
        continue;
      }
      const int col = indent + 1 + static_cast<int>(rest.size() - content.size());
      if (is_seq_item(content) || find_key_colon(content) != std::string::npos) {
        // "- key: value" / "- - x": re-interpret the remainder as a block starting at its own column
        L_[pos_].indent = col;
        L_[pos_].text = content;
        n.seq.push_back(block(col));
      } else {
        n.seq.push_back(inline_value(content));
        pos_++;
      }
    }
    return n;
  }

  Node mapping(int indent) {
    Node n;
    n.type = Node::Map;
    while (pos_ < L_.size() && L_[pos_].indent == indent) {
      const std::string& t = L_[pos_].text;
      if (is_seq_item(t)) break;
      const size_t c = find_key_colon(t);
      if (c == std::string::npos) throw std::runtime_error("yaml: expected 'key: value' at line " + std::to_string(L_[pos_].number));
      std::string key = rtrim(t.substr(0, c));
      if (key.size() >= 2 && (key[0] == '"' || key[0] == '\'') && key.back() == key[0]) key = key.substr(1, key.size() - 2);
      const std::string rest = ltrim(t.substr(c + 1));
      pos_++;
      if (!rest.empty()) {
        n.map.emplace_back(key, inline_value(rest));
      } else if (pos_ < L_.size() && L_[pos_].indent > indent) {
        n.map.emplace_back(key, block(L_[pos_].indent));
      } else if (pos_ < L_.size() && L_[pos_].indent == indent && is_seq_item(L_[pos_].text)) {
        n.map.emplace_back(key, sequence(indent));  // sequence at the key's own indentation
      } else {
        n.map.emplace_back(key, Node());
      }
    }
    if (pos_ < L_.size() && L_[pos_].indent > indent)
      throw std::runtime_error("yaml: unexpected indentation at line " + std::to_string(L_[pos_].number));
    return n;
  }
};

}  // namespace detail

inline Node Load(std::istream& in) { return detail::BlockParser(detail::logical_lines(in)).parse(); }
inline Node Load(const std::string& text) {
  std::istringstream s(text);
  return Load(s);
}
inline Node LoadFile(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("yaml: cannot open " + path);
  return Load(f);
}

}  // namespace yaml_lite
