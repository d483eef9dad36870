#ifndef RSHIM_RS_H
#define RSHIM_RS_H
#define F77_NAME(x) x##_
#define F77_CALL(x) x##_
#define F77_SUB(x) x##_
#endif
