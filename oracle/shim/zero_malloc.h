/* TEST INFRASTRUCTURE ONLY (oracle build shim) - not part of the product.
 *
 * Pre-included (-include) when compiling the reference core as the CPU oracle.
 * The reference's `backup_corr_pair` (stochqn.c:589-595) reads s_bak / y_bak, which
 * nothing ever writes; they are malloc'ed (stochqn.c:305-306) in plain C but
 * zero-filled by the R allocators (R/allocators.R:8-9).  Turning malloc into calloc
 * gives the deterministic R semantics without touching any
 * reference source.
 */
#include <stdlib.h>
#define malloc(x) calloc(1, (x))
