/* patch_globals.h -- TEST INFRASTRUCTURE: knobs for the sed-patched reference TU (oracle/Makefile). */
#pragma once
extern int rt_oracle_max_depth;  /* rays at level >= this shade as plain Phong (no child ray) */
extern int rt_oracle_usteps;     /* area-light grid, reference hard-codes 5 x 5 */
extern int rt_oracle_vsteps;
