#ifndef COMPAT_REFNL_H_
#define COMPAT_REFNL_H_
#endif
