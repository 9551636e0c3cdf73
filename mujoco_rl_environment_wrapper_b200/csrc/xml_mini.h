// xml_mini.h — minimal non-validating XML reader (elements + attributes only).
// Enough for MJCF: prolog, comments, nested elements, single/double quoted
// attributes, self-closing tags.  Text content is skipped.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mjb {

struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;

  const std::string* attr(const std::string& key) const {
    for (auto& kv : attrs)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const XmlNode* child(const std::string& t) const {
    for (auto& c : children)
      if (c->tag == t) return c.get();
    return nullptr;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& text) : s_(text), p_(0) {}

  std::unique_ptr<XmlNode> parse() {
    skip_misc();
    auto root = element();
    if (!root) throw std::runtime_error("XML: no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_;

  [[noreturn]] void fail(const std::string& msg) const {
    size_t line = 1;
    for (size_t i = 0; i < p_ && i < s_.size(); i++)
      if (s_[i] == '\n') line++;
    throw std::runtime_error("XML parse error (line " + std::to_string(line) + "): " + msg);
  }
  bool starts(const char* lit) const { return s_.compare(p_, strlen(lit), lit) == 0; }
  void skip_ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++;
  }
  // whitespace, comments, processing instructions, doctype, stray text
  void skip_misc() {
    for (;;) {
      while (p_ < s_.size() && s_[p_] != '<') p_++;
      if (p_ >= s_.size()) return;
      if (starts("<!--")) {
        size_t e = s_.find("-->", p_ + 4);
        if (e == std::string::npos) fail("unterminated comment");
        p_ = e + 3;
      } else if (starts("<?")) {
        size_t e = s_.find("?>", p_ + 2);
        if (e == std::string::npos) fail("unterminated processing instruction");
        p_ = e + 2;
      } else if (starts("<!")) {
        size_t e = s_.find('>', p_);
        if (e == std::string::npos) fail("unterminated declaration");
        p_ = e + 1;
      } else {
        return;
      }
    }
  }
  static bool name_char(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' ||
           c == '-' || c == ':' || c == '.';
  }
  std::string name() {
    size_t b = p_;
    while (p_ < s_.size() && name_char(s_[p_])) p_++;
    if (p_ == b) fail("expected a name");
    return s_.substr(b, p_ - b);
  }
  static std::string unescape(const std::string& v) {
    if (v.find('&') == std::string::npos) return v;
    std::string o;
    for (size_t i = 0; i < v.size(); i++) {
      if (v[i] == '&') {
        if (v.compare(i, 4, "&lt;") == 0) { o += '<'; i += 3; continue; }
        if (v.compare(i, 4, "&gt;") == 0) { o += '>'; i += 3; continue; }
        if (v.compare(i, 5, "&amp;") == 0) { o += '&'; i += 4; continue; }
        if (v.compare(i, 6, "&quot;") == 0) { o += '"'; i += 5; continue; }
        if (v.compare(i, 6, "&apos;") == 0) { o += '\''; i += 5; continue; }
      }
      o += v[i];
    }
    return o;
  }

  std::unique_ptr<XmlNode> element() {
    if (p_ >= s_.size() || s_[p_] != '<') return nullptr;
    p_++;
    auto node = std::make_unique<XmlNode>();
    node->tag = name();
    for (;;) {
      skip_ws();
      if (p_ >= s_.size()) fail("unterminated tag <" + node->tag + ">");
      if (s_[p_] == '/') {
        if (p_ + 1 >= s_.size() || s_[p_ + 1] != '>') fail("malformed self-closing tag");
        p_ += 2;
        return node;
      }
      if (s_[p_] == '>') { p_++; break; }
      std::string key = name();
      skip_ws();
      if (p_ >= s_.size() || s_[p_] != '=') fail("expected '=' after attribute " + key);
      p_++;
      skip_ws();
      if (p_ >= s_.size() || (s_[p_] != '"' && s_[p_] != '\'')) fail("expected quoted value for " + key);
      char q = s_[p_++];
      size_t e = s_.find(q, p_);
      if (e == std::string::npos) fail("unterminated attribute value for " + key);
      node->attrs.emplace_back(key, unescape(s_.substr(p_, e - p_)));
      p_ = e + 1;
    }
    // children until the matching close tag
    for (;;) {
      skip_misc();
      if (p_ >= s_.size()) fail("missing </" + node->tag + ">");
      if (starts("</")) {
        p_ += 2;
        std::string close = name();
        if (close != node->tag) fail("mismatched </" + close + "> for <" + node->tag + ">");
        skip_ws();
        if (p_ >= s_.size() || s_[p_] != '>') fail("malformed close tag");
        p_++;
        return node;
      }
      node->children.push_back(element());
    }
  }
};

}  // namespace mjb
