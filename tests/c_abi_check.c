/* Compiled (not run) by tests/test_abi.py with a C99 compiler: the ABI header must be plain C. */
#include "traceit_b200.h"

int use_every_entry_point(void) {
    tr_world_params p;
    tr_world* w = 0;
    tr_body_desc b;
    tr_stats st;
    tr_frame_body fr;
    uint32_t id;
    uint64_t n, a, c;
    float v[3], cell, inv;
    int32_t bits, pairs[2];
    uint32_t keys[1], ids[1];
    char uid[128];
    double ops, mhz;
    int cnt;
    (void)tr_abi_version();
    (void)tr_last_error();
    (void)tr_device_count(&cnt);
    (void)tr_default_params(&p);
    (void)tr_world_create(&p, 0, &w);
    (void)tr_comm_unique_id(uid);
    (void)tr_world_create_sharded(&p, 0, 0, 1, uid, &w);
    (void)tr_world_set_params(w, &p);
    (void)tr_world_get_params(w, &p);
    (void)tr_world_set_stream(w, 0);
    (void)tr_body_add(w, &b, &id);
    (void)tr_bodies_add(w, 1, &b, &id);
    (void)tr_world_size(w, &n);
    (void)tr_world_owned_range(w, &a, &c);
    (void)tr_to_spherical(45.0f, 0.0f, 30.0f, v);
    (void)tr_area_from_size(1.0f);
    (void)tr_step(w, 1);
    (void)tr_sync(w);
    (void)tr_step_host(w, 1, &b, 1);
    (void)tr_step_frame(w, 1, v, &fr, 1);
    (void)tr_read_state(w, 0, 1, v, v);
    (void)tr_read_bodies(w, 0, 1, &b);
    (void)tr_write_bodies(w, 0, 1, &b);
    (void)tr_set_force(w, 0, 0.f, 0.f, 0.f);
    (void)tr_world_enable_trajectory(w, 300);
    (void)tr_read_trajectory(w, 0, v, 1, &n);
    (void)tr_read_contacts(w, pairs, 1, &n);
    (void)tr_debug_grid(w, keys, ids, &n, &cell, &inv, &bits);
    (void)tr_debug_candidates(w, pairs, 1, &n);
    (void)tr_debug_gravity(w, v);
    (void)tr_get_stats(w, &st);
    (void)tr_reset_stats(w);
    (void)tr_enable_timing(w, 1);
    (void)tr_measure_fma_peak(0, &ops, &mhz);
    (void)tr_world_destroy(w);
    return (int)(sizeof(tr_body_desc) + sizeof(tr_frame_body));
}
