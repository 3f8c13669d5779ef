/* CPU ORACLE (test infrastructure, NOT product code): C restatement of the reference's Panda
 * state-/motion-validity path, in fp64 (parity margins) and fp32 (the arithmetic type of the CUDA path;
 * also the "port" CPU baseline of bench.py).  See panda_oracle_impl.h for what is restated and the
 * "parity unpinned" statement.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load the library built from this file.
 *
 * Build: make -C oracle   ->  oracle/_build/libpanda_oracle.so
 */
#include <math.h>
#include <string.h>

/* kinematic chain of the Menagerie panda.xml the reference loads at scenes.py:85 */
static const double PO_POS[11][3] = {
    {0, 0, 0}, {0, 0, 0.333}, {0, 0, 0}, {0, -0.316, 0}, {0.0825, 0, 0}, {-0.0825, 0.384, 0},
    {0, 0, 0}, {0.088, 0, 0}, {0, 0, 0.107}, {0, 0, 0.0584}, {0, 0, 0.0584}};
static const double PO_QUAT[11][4] = {
    {1, 0, 0, 0}, {1, 0, 0, 0}, {1, -1, 0, 0}, {1, 1, 0, 0}, {1, 1, 0, 0}, {1, -1, 0, 0},
    {1, 1, 0, 0}, {1, 1, 0, 0}, {0.9238795, 0, 0, -0.3826834}, {1, 0, 0, 0}, {0, 0, 0, 1}};
static const int PO_PARENT[11] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 8};
static const int PO_JTYPE[11] = {0, 1, 1, 1, 1, 1, 1, 1, 0, 2, 2}; /* 1 = revolute z, 2 = slide +y */
static const int PO_JIDX[11] = {-1, 0, 1, 2, 3, 4, 5, 6, -1, 7, 8};

#define REAL double
#define SUFFIX _f64
#include "panda_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL float
#define SUFFIX _f32
#include "panda_oracle_impl.h"
#undef REAL
#undef SUFFIX

#include "rrtc_oracle_impl.h"
