/* LD_PRELOAD helper: print a native backtrace on SIGSEGV (debugging aid, not part of the product). */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdlib.h>
#include <unistd.h>
static void handler(int sig) {
  void *bt[64];
  int n = backtrace(bt, 64);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
__attribute__((constructor)) static void init(void) { signal(SIGSEGV, handler); }
