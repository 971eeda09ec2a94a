// fill_rowtile.cu -- TXASM_SCATTER_ROWTILE (placeholder until the tile kernel lands)
#include "txasm_internal.hpp"
namespace txasm {
int tiles_build(txasm_handle h) { return set_err(h, TXASM_EUNSUPPORTED, "row-tile path not built yet"); }
void tiles_free(txasm_handle) {}
int launch_fill_rowtile(txasm_handle h, const FillArgs &) { return set_err(h, TXASM_EUNSUPPORTED, "row-tile path not built yet"); }
int tiles_info(txasm_handle, txasm_info *) { return TXASM_OK; }
}
