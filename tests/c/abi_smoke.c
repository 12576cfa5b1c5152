/* Plain-C consumer of include/ct_b200.h: proves the header is valid C (not only C++), that the
 * library links without Python, and that the host-only entry points behave.  No GPU needed:
 * ct_create must fail cleanly (CT_E_CUDA) when no device is present. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "ct_b200.h"

int main(void) {
    ct_handle h = NULL;
    ct_batch b;
    ct_idt_stage st;
    ct_idt_trace tr;
    int rc;
    memset(&b, 0, sizeof b);
    memset(&st, 0, sizeof st);
    memset(&tr, 0, sizeof tr);
    if (ct_abi_version() != CT_ABI_VERSION) return 1;
    if (sizeof(ct_batch) != 48 || sizeof(ct_idt_stage) != 96 || sizeof(ct_idt_trace) != 40) return 2;
    if (ct_idt_value_of(ct_idt_key_of(-1.5)) != -1.5 || ct_idt_key_of(-2.0) >= ct_idt_key_of(-1.0) ||
        ct_idt_key_of(-0.0) >= ct_idt_key_of(1e-300) || ct_idt_key_of(INFINITY) != 0x7ff0000000000000LL)
        return 3;
    if (ct_idt_workspace_bytes(1000, 1, 255, 4) < 3 * 1000 * sizeof(double)) return 4;
    if (CT_IDT_LUT_DOUBLES(255) != 3 * (3 * 256 + 4)) return 5;
    rc = ct_create(0, &h);
    if (rc == CT_OK) {            /* a GPU is present: the null-argument paths must still be rejected */
        if (ct_moments(h, NULL, 0, NULL) != CT_E_INVALID) return 6;
        if (ct_linear_transfer(h, 99, &b, &b, &b, NULL, NULL) != CT_E_INVALID) return 7;
        if (strlen(ct_last_error(h)) == 0) return 8;
        ct_destroy(h);
    } else if (rc != CT_E_CUDA || h != NULL) {
        return 9;
    }
    printf("abi_smoke ok (ct_create rc=%d)\n", rc);
    return 0;
}
