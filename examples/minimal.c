/* Minimal use of the C ABI: the reference's seeded world (earth slab, src/main.cpp:1482-1491) plus
 * one projectile built the way its creation form does (src/main.cpp:1293-1307), stepped 600 frames
 * on the B200 with the trajectory history on.
 *   gcc -std=c99 -Iinclude examples/minimal.c -Ltraceit_b200 -ltraceit_b200 -Wl,-rpath,$PWD/traceit_b200 -o minimal */
#include <stdio.h>
#include <string.h>

#include "traceit_b200.h"

#define CHECK(call) do { if ((call) != TR_OK) { fprintf(stderr, "%s failed: %s\n", #call, tr_last_error()); return 1; } } while (0)

int main(void) {
    tr_world_params p;
    tr_world* w = NULL;
    tr_body_desc b[2];
    float pos[6], vel[6], traj[3 * 300];
    uint64_t npts = 0;
    int frame;

    CHECK(tr_default_params(&p));            /* g, dt, atmDensity = the reference's constants */
    CHECK(tr_world_create(&p, 0, &w));
    CHECK(tr_world_enable_trajectory(w, 300));

    memset(b, 0, sizeof(b));
    /* earth */
    b[0].pos[1] = b[0].center[1] = -20.0f;
    b[0].mass = 1000.0f; b[0].Cd = 1000.0f; b[0].area = 1000.0f; b[0].stationary = 1;
    b[0].size[0] = 10000.0f; b[0].size[1] = 5.0f; b[0].size[2] = 10000.0f; b[0].rest = 1.0f;
    /* projectile: mass 10, elevation 45 deg, azimuth 0, magnitude 30, size 1 */
    b[1].mass = 10.0f; b[1].Cd = 0.42f; b[1].rest = 0.5f;
    CHECK(tr_to_spherical(45.0f, 0.0f, 30.0f, b[1].vel));
    b[1].area = tr_area_from_size(1.0f);
    b[1].size[0] = b[1].size[1] = b[1].size[2] = 1.0f;
    CHECK(tr_bodies_add(w, 2, b, NULL));

    for (frame = 0; frame < 600; ++frame) CHECK(tr_step(w, 1));   /* updatePhysics(w, coords) */
    CHECK(tr_sync(w));
    CHECK(tr_read_state(w, 0, 2, pos, vel));
    CHECK(tr_read_trajectory(w, 1, traj, 300, &npts));
    printf("projectile after 600 frames: pos (%g, %g, %g) vel (%g, %g, %g); %llu trajectory points, oldest (%g, %g, %g)\n",
           pos[3], pos[4], pos[5], vel[3], vel[4], vel[5], (unsigned long long)npts, traj[0], traj[1], traj[2]);
    CHECK(tr_world_destroy(w));
    return 0;
}
