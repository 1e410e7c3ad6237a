/* oracle/_ref build only: defines posix_memalign inside the reference plugin's own shared object
 * (linked -Bsymbolic) so that the reference's scratch pool (SangNom2.cpp:290-306 calls
 * posix_memalign directly and never clears the result) starts zero-filled. That is the parity
 * contract stated in DESIGN.md ("each frame as by a freshly constructed instance whose pool is
 * zero-filled"); without it the never-written pool rows hold whatever the heap held. */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <errno.h>

int posix_memalign(void** out, size_t align, size_t size)
{
    size_t rounded = (size + align - 1) / align * align;
    void* p = aligned_alloc(align, rounded ? rounded : align);
    if (!p) return ENOMEM;
    memset(p, 0, rounded ? rounded : align);
    *out = p;
    return 0;
}
