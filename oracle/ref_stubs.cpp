// TEST INFRASTRUCTURE (oracle).  Not part of the product.
//
// Link-time stand-ins used only when the reference at /root/reference is built
// headless into oracle/_ref/ (see oracle/Makefile).  CGL::OSDText is a member
// of the reference's Application object (src/application/application.h) and its
// real implementation needs freetype + a GL context; the `-f` render-to-file
// path (src/application/main.cpp:169-182) never draws text, so empty bodies are
// enough.  Nothing here restates or replaces reference arithmetic.
#include "CGL/osdtext.h"

namespace CGL {
OSDText::OSDText() {}
OSDText::~OSDText() {}
int OSDText::init(bool) { return 0; }
void OSDText::render() {}
void OSDText::clear() {}
void OSDText::resize(size_t, size_t) {}
int OSDText::add_line(float, float, std::string, size_t, Color) { return 0; }
}  // namespace CGL
