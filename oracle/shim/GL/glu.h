/* Minimal stand-in for <GL/glu.h>: glew.h includes it, and the reference calls
 * only these two entry points (application.cpp).  They are never reached on the
 * headless -f path; the symbols stay unresolved at link time. */
#ifndef ORACLE_SHIM_GLU_H
#define ORACLE_SHIM_GLU_H
#ifdef __cplusplus
extern "C" {
#endif
void gluLookAt(double, double, double, double, double, double, double, double, double);
void gluPerspective(double, double, double, double);
#ifdef __cplusplus
}
#endif
#endif
