/* ORACLE INFRASTRUCTURE: empty stand-ins for the 8 OpenGL symbols the reference test binary imports (drawing only). */
void glBegin(int a) { (void)a; } void glCallList(int a) { (void)a; } void glEnd(void) {} void glEndList(void) {}
unsigned glGenLists(int n) { (void)n; return 1; } void glNewList(unsigned l, int m) { (void)l; (void)m; }
void glNormal3dv(const double* v) { (void)v; } void glVertex3dv(const double* v) { (void)v; }
