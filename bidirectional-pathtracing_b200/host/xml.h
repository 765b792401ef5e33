// xml.h — a small DOM parser for the subset of XML that COLLADA files use: elements, attributes,
// character data, comments, processing instructions, CDATA, the five predefined entities.
// (The reference parses with a vendored tinyxml2; this host has no third-party code.)
#pragma once
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace bdpt_host {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;
  std::string text;  // character data that precedes the first child element (what tinyxml2's GetText() returns)
  bool has_text = false;
  XmlNode* parent = nullptr;

  const char* attr(const char* key) const {
    for (const auto& a : attrs) if (a.first == key) return a.second.c_str();
    return nullptr;
  }
  XmlNode* first_child(const char* nm = nullptr) const {
    for (const auto& c : children) if (!nm || c->name == nm) return c.get();
    return nullptr;
  }
  XmlNode* next_sibling(const char* nm = nullptr) const {
    if (!parent) return nullptr;
    bool seen = false;
    for (const auto& c : parent->children) {
      if (seen && (!nm || c->name == nm)) return c.get();
      if (c.get() == this) seen = true;
    }
    return nullptr;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& s) : s_(s), p_(0) {}

  std::unique_ptr<XmlNode> parse() {
    skip_misc();
    std::unique_ptr<XmlNode> root = element(nullptr);
    if (!root) throw std::runtime_error("XML: no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_;

  bool starts(const char* lit) const { return s_.compare(p_, strlen(lit), lit) == 0; }
  void skip_ws() { while (p_ < s_.size() && (unsigned char)s_[p_] <= ' ') p_++; }
  void skip_until(const char* end) {
    size_t q = s_.find(end, p_);
    if (q == std::string::npos) throw std::runtime_error("XML: unterminated construct");
    p_ = q + strlen(end);
  }
  void skip_misc() {  // whitespace, BOM, <?...?>, <!--...-->, <!DOCTYPE ...>
    if (s_.compare(0, 3, "\xEF\xBB\xBF") == 0 && p_ == 0) p_ = 3;
    for (;;) {
      skip_ws();
      if (starts("<?")) skip_until("?>");
      else if (starts("<!--")) skip_until("-->");
      else if (starts("<!DOCTYPE")) skip_until(">");
      else return;
    }
  }
  static void append_decoded(std::string& out, const std::string& raw) {
    for (size_t i = 0; i < raw.size(); i++) {
      if (raw[i] != '&') { out += raw[i]; continue; }
      size_t e = raw.find(';', i);
      if (e == std::string::npos) { out += raw[i]; continue; }
      std::string ent = raw.substr(i + 1, e - i - 1);
      if (ent == "amp") out += '&';
      else if (ent == "lt") out += '<';
      else if (ent == "gt") out += '>';
      else if (ent == "quot") out += '"';
      else if (ent == "apos") out += '\'';
      else if (!ent.empty() && ent[0] == '#') {
        long c = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X') ? strtol(ent.c_str() + 2, nullptr, 16) : strtol(ent.c_str() + 1, nullptr, 10);
        if (c > 0 && c < 128) out += (char)c;
      } else { out += raw.substr(i, e - i + 1); }
      i = e;
    }
  }
  std::string name() {
    size_t b = p_;
    while (p_ < s_.size()) {
      char c = s_[p_];
      if ((unsigned char)c <= ' ' || c == '>' || c == '/' || c == '=') break;
      p_++;
    }
    return s_.substr(b, p_ - b);
  }

  std::unique_ptr<XmlNode> element(XmlNode* parent) {
    if (p_ >= s_.size() || s_[p_] != '<') return nullptr;
    p_++;
    std::unique_ptr<XmlNode> n(new XmlNode());
    n->parent = parent;
    n->name = name();
    for (;;) {  // attributes
      skip_ws();
      if (p_ >= s_.size()) throw std::runtime_error("XML: unexpected end in tag");
      if (s_[p_] == '/') { skip_until(">"); return n; }
      if (s_[p_] == '>') { p_++; break; }
      std::string key = name();
      skip_ws();
      if (p_ >= s_.size() || s_[p_] != '=') throw std::runtime_error("XML: attribute without value");
      p_++;
      skip_ws();
      char q = s_[p_];
      if (q != '"' && q != '\'') throw std::runtime_error("XML: unquoted attribute");
      size_t e = s_.find(q, p_ + 1);
      if (e == std::string::npos) throw std::runtime_error("XML: unterminated attribute");
      std::string val;
      append_decoded(val, s_.substr(p_ + 1, e - p_ - 1));
      n->attrs.emplace_back(key, val);
      p_ = e + 1;
    }
    for (;;) {  // content
      size_t lt = s_.find('<', p_);
      if (lt == std::string::npos) throw std::runtime_error("XML: missing end tag of " + n->name);
      if (lt > p_ && n->children.empty()) {
        std::string raw = s_.substr(p_, lt - p_);
        bool blank = true;
        for (char c : raw) if ((unsigned char)c > ' ') { blank = false; break; }
        if (!blank || n->has_text) { append_decoded(n->text, raw); n->has_text = true; }
      }
      p_ = lt;
      if (starts("</")) { skip_until(">"); return n; }
      if (starts("<!--")) { skip_until("-->"); continue; }
      if (starts("<![CDATA[")) {
        size_t e = s_.find("]]>", p_);
        if (e == std::string::npos) throw std::runtime_error("XML: unterminated CDATA");
        if (n->children.empty()) { n->text += s_.substr(p_ + 9, e - p_ - 9); n->has_text = true; }
        p_ = e + 3;
        continue;
      }
      if (starts("<?")) { skip_until("?>"); continue; }
      n->children.push_back(element(n.get()));
    }
  }
};

}  // namespace bdpt_host
