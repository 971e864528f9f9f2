// floor1_emu.cpp -- TEST INFRASTRUCTURE: the floor-1 evaluation source the GPU compiles
// (lewton_b200/csrc/floor1_eval.cuh + prepare_floor1 of tables_host.cpp), run on the host:
// post unwrap, closed-form curve and the DDA segment render, for the oracle to check.
#include <cstring>
#include <vector>

#include "../../lewton_b200/csrc/floor1_eval.cuh"
#include "../../lewton_b200/csrc/tables_host.cpp"

extern "C" int lwb_emu_floor1(int mult, const uint32_t *xs, int nposts, const uint32_t *y, int n2,
                              uint32_t *curve_closed, uint8_t *curve_render)
{
    lwb_floor_desc d;
    std::memset(&d, 0, sizeof(d));
    d.floor_type = LWB_FLOOR_TYPE_ONE;
    d.floor1_multiplier = (uint8_t)mult;
    d.floor1_values = (uint8_t)nposts;
    for (int i = 0; i < nposts; i++) d.floor1_x_list[i] = xs[i];
    lwb::DevFloor1 fl;
    const int rc = lwb::prepare_floor1(d, &fl);
    if (rc) return rc;
    uint16_t sx[LWB_MAX_POSTS + 1], sy[LWB_MAX_POSTS + 1];
    const int m = lwb::d_floor1_posts(fl, y, n2, sx, sy);
    for (int k = 0; k < n2; k++) curve_closed[k] = lwb::d_floor1_y_at(sx, sy, m, k);
    std::memset(curve_render, 0xee, (size_t)n2);
    for (int seg = 0; seg + 1 < m; seg++) lwb::d_floor1_render_segment(sx, sy, seg, n2, curve_render);
    return 0;
}
