/* Plain-C consumer of include/b200krylov.h: proves that the header is C-clean (compiled with gcc -std=c99 -pedantic,
 * no C++), that the option structs have the sizes the library reports, and that the host-side entry points can be
 * called from C.  No device is needed: only b200k_version / b200k_sizeof / *_opts_default / b200k_exponential /
 * b200k_expv_small / b200k_status_string run. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "b200krylov.h"

int main(void) {
    b200k_krylov_opts ko;
    b200k_kiops_opts io;
    b200k_timestep_opts to;
    double A[4] = {0.0, 0.0, 1.0, 0.0}; /* column-major [[0, 1], [0, 0]]: exp(A) = [[1, 1], [0, 1]] */
    double H[9] = {-2.0, 1.0, 0.0, 1.0, -2.0, 1.0, 0.0, 1.0, -2.0}; /* symmetric tridiagonal */
    double y[3];
    int branch = -1;

    if (b200k_version() != B200K_VERSION) return 1;
    if (b200k_sizeof(B200K_STRUCT_KRYLOV_OPTS) != (int)sizeof(b200k_krylov_opts)) return 2;
    if (b200k_sizeof(B200K_STRUCT_KIOPS_OPTS) != (int)sizeof(b200k_kiops_opts)) return 3;
    if (b200k_sizeof(B200K_STRUCT_TIMESTEP_OPTS) != (int)sizeof(b200k_timestep_opts)) return 4;
    if (b200k_sizeof(99) != -1) return 5;

    b200k_krylov_opts_default(&ko);
    b200k_kiops_opts_default(&io);
    b200k_timestep_opts_default(&to);
    if (ko.m != 30 || ko.tol != 1.0e-7 || ko.iop != 0 || ko.hermitian != -1 || ko.init != 0 || ko.p != 0) return 6;
    if (io.mmin != 10 || io.mmax != 128 || io.m != 10 || io.iop != 2 || io.task1 != 0 || io.normU == io.normU) return 7;
    if (to.m != 10 || to.tol != 1.0e-7 || to.delta != 1.2 || to.gamma != 0.8 || to.adaptive != 0) return 8;

    if (b200k_exponential(2, A, 2) != B200K_OK) return 9;
    if (fabs(A[0] - 1.0) > 1e-15 || fabs(A[1]) > 1e-15 || fabs(A[2] - 1.0) > 1e-15 || fabs(A[3] - 1.0) > 1e-15) return 10;

    if (b200k_expv_small(3, H, 3, 0.5, y, &branch) != B200K_OK) return 11;
    if (branch != 1) return 12; /* exactly symmetric -> SymTridiagonal eigen branch (krylov_phiv.jl:225-229) */
    if (!(y[0] > 0.0 && y[0] < 1.0 && y[1] > 0.0 && y[2] > 0.0)) return 13;

    if (strcmp(b200k_status_string(B200K_OK), "ok") != 0) return 14;
    printf("consumer ok: version %d, sizes %d/%d/%d, y = %.17g %.17g %.17g\n", b200k_version(),
           b200k_sizeof(1), b200k_sizeof(2), b200k_sizeof(3), y[0], y[1], y[2]);
    return 0;
}
