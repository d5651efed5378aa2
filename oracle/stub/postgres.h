/* Minimal stand-in for PostgreSQL's postgres.h, sufficient for /root/reference/distfunc.c
 * (which only needs bool / size_t / fixed-width ints from it; distfunc.c:15).
 * TEST INFRASTRUCTURE ONLY -- used to compile the unmodified reference into oracle/_ref/. */
#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>
