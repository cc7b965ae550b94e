// Shim over the reference's XML parser, pugixml as vendored under /root/reference/ext/pugixml, compiled where it lies
// (oracle/ref/Makefile).  TEST INFRASTRUCTURE: dumps the element tree of a scene file in a canonical text form
// ("<depth> <name> <attr>=<value> ...\n", document order) so that host/XmlLite.h can be compared with it.
#include <cstdint>
#include <cstring>
#include <string>
#include "pugixml.cpp"   // the reference's copy (-I$(REF)/ext/pugixml/src)

static void dump(const pugi::xml_node& n, int depth, std::string& out) {
	out += std::to_string(depth) + " " + n.name();
	for (const pugi::xml_attribute& a : n.attributes()) out += std::string(" ") + a.name() + "=" + a.value();
	out += "\n";
	for (const pugi::xml_node& c : n.children()) if (c.type() == pugi::node_element) dump(c, depth + 1, out);
}

// returns the number of bytes the dump needs (including the terminator); writes at most `capacity` bytes
extern "C" __attribute__((visibility("default")))
size_t ref_xml_dump(const char* path, char* out, size_t capacity) {
	pugi::xml_document doc;
	if (!doc.load_file(path)) return 0;   // the reference's call: src/Scene.cpp:107
	std::string s;
	for (const pugi::xml_node& c : doc.children()) if (c.type() == pugi::node_element) dump(c, 0, s);
	if (out && capacity) { std::strncpy(out, s.c_str(), capacity - 1); out[capacity - 1] = 0; }
	return s.size() + 1;
}
