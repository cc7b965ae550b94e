// Minimal XML reader for the scene dialect the reference parses with pugixml (src/Scene.cpp:104-190,
// src/Material.cpp).  Elements, attributes, comments, <?xml ?> prolog; no entities beyond the basic five,
// no CDATA, no namespaces — the scene files use none of those.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rpt {

struct XmlNode {
	std::string name;
	std::vector<std::pair<std::string, std::string>> attrs;
	std::vector<std::unique_ptr<XmlNode>> children;

	const XmlNode* child(const std::string& n) const {
		for (auto& c : children) if (c->name == n) return c.get();
		return nullptr;
	}
	bool hasAttr(const std::string& n) const {
		for (auto& a : attrs) if (a.first == n) return true;
		return false;
	}
	// pugixml's as_string() of a missing attribute is ""
	std::string attr(const std::string& n) const {
		for (auto& a : attrs) if (a.first == n) return a.second;
		return "";
	}
};

class XmlParser {
public:
	explicit XmlParser(const std::string& text) : s(text) {}

	std::unique_ptr<XmlNode> parseDocument() {
		skipMisc();
		auto root = parseElement();
		if (!root) throw std::runtime_error("XmlLite: no root element");
		return root;
	}

private:
	const std::string& s;
	size_t p = 0;

	bool startsWith(const char* t) const { return s.compare(p, std::char_traits<char>::length(t), t) == 0; }
	void skipWs() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++; }

	void skipMisc() {
		for (;;) {
			skipWs();
			if (startsWith("<?")) { size_t e = s.find("?>", p); p = (e == std::string::npos) ? s.size() : e + 2; }
			else if (startsWith("<!--")) { size_t e = s.find("-->", p); p = (e == std::string::npos) ? s.size() : e + 3; }
			else if (startsWith("<!")) { size_t e = s.find('>', p); p = (e == std::string::npos) ? s.size() : e + 1; }
			else break;
		}
	}

	std::string parseName() {
		size_t b = p;
		while (p < s.size()) {
			char c = s[p];
			if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '=' || c == '>' || c == '/') break;
			p++;
		}
		return s.substr(b, p - b);
	}

	static std::string unescape(const std::string& v) {
		std::string o;
		for (size_t i = 0; i < v.size(); i++) {
			if (v[i] == '&') {
				static const char* ent[][2] = { {"&lt;", "<"}, {"&gt;", ">"}, {"&amp;", "&"}, {"&quot;", "\""}, {"&apos;", "'"} };
				bool hit = false;
				for (auto& e : ent) {
					size_t n = std::char_traits<char>::length(e[0]);
					if (v.compare(i, n, e[0]) == 0) { o += e[1]; i += n - 1; hit = true; break; }
				}
				if (!hit) o += v[i];
			}
			else o += v[i];
		}
		return o;
	}

	std::unique_ptr<XmlNode> parseElement() {
		skipMisc();
		if (p >= s.size() || s[p] != '<') return nullptr;
		p++;
		auto node = std::make_unique<XmlNode>();
		node->name = parseName();
		for (;;) {
			skipWs();
			if (p >= s.size()) throw std::runtime_error("XmlLite: unexpected end in <" + node->name);
			if (s[p] == '/') { p += 2; return node; }            // "/>"
			if (s[p] == '>') { p++; break; }
			std::string key = parseName();
			skipWs();
			if (p >= s.size() || s[p] != '=') throw std::runtime_error("XmlLite: expected '=' after " + key);
			p++;
			skipWs();
			char q = s[p++];
			size_t e = s.find(q, p);
			if (e == std::string::npos) throw std::runtime_error("XmlLite: unterminated attribute " + key);
			node->attrs.emplace_back(key, unescape(s.substr(p, e - p)));
			p = e + 1;
		}
		for (;;) {
			// skip text content
			while (p < s.size() && s[p] != '<') p++;
			if (p >= s.size()) throw std::runtime_error("XmlLite: missing </" + node->name + ">");
			if (startsWith("</")) {
				size_t e = s.find('>', p);
				p = (e == std::string::npos) ? s.size() : e + 1;
				return node;
			}
			if (startsWith("<!--") || startsWith("<?") || startsWith("<!")) { skipMisc(); continue; }
			node->children.push_back(parseElement());
		}
	}
};

} // namespace rpt
