// session.inl -- resident sessions (include/rebx_b200.h) and the operator API (include/reboundx.h), included at the end
// of dispatch.cpp: both need the dispatcher's table cache, effect entries and device context.
//
// A session keeps a simulation's particles as SoA state in HBM between force evaluations and drives the device layer
// from C: sweep, fused drift-kick-drift step, Wisdom-Holman step, integrate_force schemes.  The operator entry points
// (rebx_load_operator("integrate_force") & co., reference src/core.c:362-453,500-602, src/integrate_force.c:34-78,
// src/steppers.c:109-130) are the reference's own plug-in surface for the same work; rebx_integrate_force runs on a
// short-lived session when its force is device-backed and on the host schemes otherwise.

struct rebx_b200_session {
    struct reb_simulation* sim = nullptr;
    struct rebx_extras* rebx = nullptr;
    rebxh::Side* side = nullptr;
    rebxh::DeviceContext* ctx = nullptr;
    size_t N = 0, stride = 0;
    rebxh::DevBuf aos, soa, bodies, br, work, scratch;
    std::vector<rebxh::Entry> entries;
    rebx_b200_pass P;
    std::uint64_t generation = 0, epoch = 0;
    bool reduces = false, fusable_wh = false;
    double stepped = 0.0;           // time advanced since the last sync
    int launches_i = 0;
    std::uint64_t launches = 0;
    cudaStream_t stream = nullptr;
};

namespace rebxh {

static const rebx_b200_aos_layout kLayout = {(int32_t)sizeof(struct reb_particle), (int32_t)offsetof(struct reb_particle, x),
                                             (int32_t)offsetof(struct reb_particle, vx), (int32_t)offsetof(struct reb_particle, ax),
                                             (int32_t)offsetof(struct reb_particle, m), (int32_t)offsetof(struct reb_particle, r)};

static int session_fail(rebx_b200_session* s, const char* what, int code) {
    char buf[400];
    snprintf(buf, sizeof buf, "REBOUNDx-B200 Error: resident session: %s%s%s.\n", what, code > 0 ? ": " : "", code > 0 ? cudaGetErrorString((cudaError_t)code) : "");
    if (s && s->sim) reb_simulation_error(s->sim, buf); else fprintf(stderr, "%s", buf);
    cudaGetLastError();
    return code ? code : -1;
}

// (Re)builds the effect entries, uploads tables and body records, fills the pass descriptor.
static int session_bind(rebx_b200_session* s, struct rebx_force** only, int n_only) {
    struct reb_simulation* sim = s->sim;
    s->entries.clear();
    std::vector<struct rebx_force*> order;
    if (only) order.assign(only, only + n_only);
    else for (struct rebx_node* n = s->rebx->additional_forces; n; n = n->next) order.push_back(static_cast<struct rebx_force*>(n->object));
    if (order.empty()) return session_fail(s, "no force has been added to the simulation", 0);
    for (struct rebx_force* f : order) {
        const int kind = kind_of(f);
        if (kind == REBX_B200_NONE || kind == KIND_GR) return session_fail(s, "every force of a session must be a device-backed sweep (custom forces and `gr` run through sim->additional_forces)", 0);
        entries_for_force(sim, s->side, s->ctx, f, kind, sim->particles, s->N, s->entries, s->ctx->tables);
    }
    if (s->entries.empty() || s->entries.size() > REBX_B200_MAX_EFFECTS) return session_fail(s, "a session holds 1..4 effect entries", 0);
    for (const Entry& e : s->entries) if (e.frame != 0 || e.body_index != 0) return session_fail(s, "every effect of a session must refer to particle 0 in PARTICLE coordinates", 0);
    const size_t ne = s->entries.size();
    cudaError_t err;
    if ((err = s->bodies.ensure(ne*REBX_B200_BODY_DOUBLES*sizeof(double))) != cudaSuccess) return session_fail(s, "allocating body records", err);
    if ((err = s->br.ensure(ne*3*sizeof(double))) != cudaSuccess) return session_fail(s, "allocating back-reaction sums", err);
    std::vector<double> rec(ne*REBX_B200_BODY_DOUBLES);
    for (size_t e = 0; e < ne; e++) fill_body(rec.data() + e*REBX_B200_BODY_DOUBLES, s->entries[e], sim->particles);
    if ((err = cudaMemcpyAsync(s->bodies.p, rec.data(), rec.size()*sizeof(double), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) return session_fail(s, "copying body records", err);
    if ((err = cudaStreamSynchronize(s->stream)) != cudaSuccess) return session_fail(s, "copying body records", err);
    std::memset(&s->P, 0, sizeof s->P);
    double* base = static_cast<double*>(s->soa.p);
    rebx_b200_state& st = s->P.st;
    st.n = (int64_t)s->N; st.index_base = 0;
    st.x = base; st.y = base + s->stride; st.z = base + 2*s->stride;
    st.vx = base + 3*s->stride; st.vy = base + 4*s->stride; st.vz = base + 5*s->stride;
    st.m = base + 6*s->stride; st.r = base + 7*s->stride;
    st.ax = base + 8*s->stride; st.ay = base + 9*s->stride; st.az = base + 10*s->stride;
    s->P.n_effects = (int32_t)ne;
    s->P.reserved = REBX_B200_PASS_SHARED_BODY | (s->ctx->tolerance ? REBX_B200_PASS_TOLERANCE : 0);
    s->reduces = false;
    for (size_t j = 0; j < ne; j++) {
        Entry& en = s->entries[j];
        if (en.T && (err = upload_tables(*en.T, s->ctx->stats, 0, 0, s->N)) != cudaSuccess) return session_fail(s, "uploading parameter tables", err);
        rebx_b200_effect& fx = s->P.fx[j];
        fx.kind = en.kind; fx.flags = en.flags;
        fx.body = static_cast<const double*>(s->bodies.p) + j*REBX_B200_BODY_DOUBLES;
        for (int k = 0; k < REBX_B200_MAX_SCALARS; k++) fx.scal[k] = en.scal[k];
        if (en.T) {
            const ForceTables::Slice& S = *en.T->dev[0];
            for (int k = 0; k < REBX_B200_MAX_TABLES; k++) fx.tab[k] = static_cast<const double*>(S.d[k].p);
            fx.itab = static_cast<const int32_t*>(S.di.p);
        }
        fx.backreaction = en.reduces ? static_cast<double*>(s->br.p) + j*3 : nullptr;
        if (en.reduces) s->reduces = true;
    }
    s->fusable_wh = ne == 1 && !s->reduces && (s->entries[0].kind == REBX_B200_RADIATION || s->entries[0].kind == REBX_B200_YARKOVSKY);
    s->generation = s->side->generation;
    s->epoch = global_epoch();
    return 0;
}

// Parameters changed through the API since the tables were built: rebuild (cheap check per call).
static int session_refresh(rebx_b200_session* s) {
    refresh_escaped(s->side);
    if (s->generation == s->side->generation && s->epoch == global_epoch()) return 0;
    return session_bind(s, nullptr, 0);
}

#define REBXS(call, what) do { const int rc__ = (int)(call); if (rc__ != 0) return session_fail(s, what, rc__); } while (0)

static rebx_b200_session* session_open(struct reb_simulation* sim, struct rebx_extras* rebx, struct rebx_force** only, int n_only) {
    if (!sim || !rebx || sim->N < 2) { fprintf(stderr, "REBOUNDx-B200 Error: resident session needs a simulation with at least two particles.\n"); return nullptr; }
    Side* side = side_of(rebx);
    std::string why;
    DeviceContext* ctx = ensure_ctx(side, why);
    if (!ctx) {
        std::string msg = "REBOUNDx-B200 Error: " + why + ". Resident sessions run on the GPU only.\n";
        reb_simulation_error(sim, msg.c_str());
        return nullptr;
    }
    auto s = std::make_unique<rebx_b200_session>();
    s->sim = sim; s->rebx = rebx; s->side = side; s->ctx = ctx;
    s->N = sim->N; s->stride = (s->N + 1) & ~size_t(1);
    s->stream = ctx->devs[0]->stream[0];
    cudaError_t err;
    if ((err = s->aos.ensure(s->N*sizeof(struct reb_particle))) != cudaSuccess || (err = s->soa.ensure(11*s->stride*sizeof(double))) != cudaSuccess ||
        (err = s->work.ensure_zeroed(rebx_b200_workspace_bytes())) != cudaSuccess) { session_fail(s.get(), "allocating device state", err); return nullptr; }
    if ((err = copy_page_split(s->aos.p, sim->particles, s->N*sizeof(struct reb_particle), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess) { session_fail(s.get(), "copying particles to the device", err); return nullptr; }
    ctx->stats.h2d_bytes += s->N*sizeof(struct reb_particle);
    if (session_bind(s.get(), only, n_only) != 0) return nullptr;
    if (rebx_b200_unpack_aos(s->aos.p, &kLayout, &s->P.st, s->stream, &s->launches_i) != 0) { session_fail(s.get(), "unpacking particles", 0); return nullptr; }
    return s.release();
}

static int session_gather(rebx_b200_session* s) {
    return rebx_b200_gather_bodies(&s->P.st, static_cast<double*>(s->bodies.p), s->P.n_effects, s->stream, &s->launches_i);
}

static int session_force(rebx_b200_session* s) {        // a += stacked effects (+ back-reaction onto particle 0)
    rebx_b200_pass Q = s->P;                             // one shard holds everything: the sweep's last CTA applies the sums
    Q.reserved |= REBX_B200_PASS_APPLY_BACKREACTION;
    REBXS(rebx_b200_launch_pass(&Q, s->work.p, s->stream, &s->launches_i), "launching the force sweep");
    return 0;
}

static void session_count(rebx_b200_session* s) { s->launches += (std::uint64_t)s->launches_i; s->ctx->stats.launches += (std::uint64_t)s->launches_i; s->launches_i = 0; }

}  // namespace rebxh

extern "C" {

rebx_b200_session* rebx_b200_session_create(struct reb_simulation* sim, struct rebx_extras* rebx) {
    return rebxh::session_open(sim, rebx, nullptr, 0);
}

void rebx_b200_session_free(rebx_b200_session* s) {
    if (!s) return;
    if (s->stream) cudaStreamSynchronize(s->stream);
    delete s;
}

uint64_t rebx_b200_session_launches(const rebx_b200_session* s) { return s ? s->launches + (uint64_t)s->launches_i : 0; }

int rebx_b200_session_evaluate(rebx_b200_session* s) {
    using namespace rebxh;
    if (!s) return -1;
    REBXS(session_refresh(s), "refreshing the parameter tables");
    REBXS(session_gather(s), "refreshing the body records");
    const int rc = session_force(s);
    session_count(s);
    return rc;
}

int rebx_b200_session_step_leapfrog(rebx_b200_session* s, int nsteps, double dt) {
    using namespace rebxh;
    if (!s) return -1;
    REBXS(session_refresh(s), "refreshing the parameter tables");
    const double G = s->sim->G, soft = s->sim->softening;
    REBXS(session_gather(s), "refreshing the body records");
    for (int k = 0; k < nsteps; k++) {
        // one fused sweep per step where the stack has an instantiation; else the same arithmetic in separate sweeps
        int rc = rebx_b200_leapfrog_step(&s->P, dt, G, soft, s->work.p, s->stream, &s->launches_i);
        if (rc == cudaErrorNotSupported) {
            REBXS(rebx_b200_drift(&s->P.st, 0.5*dt, s->stream, &s->launches_i), "drift");
            REBXS(session_gather(s), "refreshing the body records");
            REBXS(rebx_b200_point_mass_gravity(&s->P.st, static_cast<const double*>(s->bodies.p), 1, G, soft, s->stream, &s->launches_i), "central gravity");
            REBXS(session_force(s), "force");
            REBXS(rebx_b200_kick(&s->P.st, dt, s->stream, &s->launches_i), "kick");
            REBXS(rebx_b200_drift(&s->P.st, 0.5*dt, s->stream, &s->launches_i), "drift");
            REBXS(session_gather(s), "refreshing the body records");
        } else if (rc != 0) return session_fail(s, "fused drift-kick-drift step", rc);
    }
    s->stepped += nsteps*dt;
    session_count(s);
    return 0;
}

int rebx_b200_session_step_wh(rebx_b200_session* s, int nsteps, double dt) {
    using namespace rebxh;
    if (!s) return -1;
    if (nsteps <= 0) return 0;
    REBXS(session_refresh(s), "refreshing the parameter tables");
    const double G = s->sim->G;
    const double* body = static_cast<const double*>(s->bodies.p);
    REBXS(session_gather(s), "refreshing the body records");
    REBXS(rebx_b200_kepler_drift(&s->P.st, body, G, 0.5*dt, s->stream, &s->launches_i), "Kepler drift");
    REBXS(session_gather(s), "refreshing the body records");
    for (int k = 0; k < nsteps; k++) {
        const double drift = (k + 1 < nsteps) ? dt : 0.5*dt;        // consecutive half drifts merged, as WHFast does between outputs
        if (s->fusable_wh) {
            REBXS(rebx_b200_wh_kick_drift(&s->P, G, dt, drift, s->stream, &s->launches_i), "fused kick + Kepler drift");
        } else {
            REBXS(rebx_b200_zero_acc(&s->P.st, s->stream, &s->launches_i), "clearing accelerations");
            REBXS(session_force(s), "force");
            REBXS(rebx_b200_kick(&s->P.st, dt, s->stream, &s->launches_i), "kick");
            REBXS(session_gather(s), "refreshing the body records");
            REBXS(rebx_b200_kepler_drift(&s->P.st, body, G, drift, s->stream, &s->launches_i), "Kepler drift");
        }
        REBXS(session_gather(s), "refreshing the body records");
    }
    s->stepped += nsteps*dt;
    session_count(s);
    return 0;
}

int rebx_b200_session_integrate_force(rebx_b200_session* s, double dt, int scheme) {
    using namespace rebxh;
    if (!s) return -1;
    REBXS(session_refresh(s), "refreshing the parameter tables");
    REBXS(s->scratch.ensure(12*s->N*sizeof(double)), "allocating stage arrays");
    REBXS(session_gather(s), "refreshing the body records");
    int iterations = -1;
    REBXS(rebx_b200_integrate_force(&s->P, dt, scheme, s->scratch.p, s->work.p, s->stream, &s->launches_i, &iterations), "integrate_force");
    if (scheme == 0 && iterations == 10)
        reb_simulation_warning(s->sim, "REBOUNDx: 10 iterations in integrator_implicit_midpoint.c failed to converge. This is typically because the perturbation is too strong for the current implementation.");
    session_count(s);
    return 0;
}

int rebx_b200_session_sync_to_host(rebx_b200_session* s) {
    using namespace rebxh;
    if (!s) return -1;
    // positions, velocities and accelerations back into the AoS records (three strided scatters), one DMA to the host
    const rebx_b200_state& st = s->P.st;
    rebx_b200_aos_layout L = kLayout;
    rebx_b200_state v = st;
    L.off_a = kLayout.off_x; v.ax = const_cast<double*>(st.x); v.ay = const_cast<double*>(st.y); v.az = const_cast<double*>(st.z);
    REBXS(rebx_b200_pack_acc_aos(s->aos.p, &L, &v, s->stream, &s->launches_i), "packing positions");
    L.off_a = kLayout.off_v; v.ax = const_cast<double*>(st.vx); v.ay = const_cast<double*>(st.vy); v.az = const_cast<double*>(st.vz);
    REBXS(rebx_b200_pack_acc_aos(s->aos.p, &L, &v, s->stream, &s->launches_i), "packing velocities");
    REBXS(rebx_b200_pack_acc_aos(s->aos.p, &kLayout, &st, s->stream, &s->launches_i), "packing accelerations");
    REBXS(copy_page_split(s->sim->particles, s->aos.p, s->N*sizeof(struct reb_particle), cudaMemcpyDeviceToHost, s->stream), "copying the state to the host");
    REBXS(cudaStreamSynchronize(s->stream), "waiting for the device");
    s->ctx->stats.d2h_bytes += s->N*sizeof(struct reb_particle);
    s->sim->t += s->stepped;
    s->stepped = 0.0;
    session_count(s);
    return 0;
}

// ---- operator API (reference src/core.c:362-453, 500-602, 1041-1089) -------------------------------------------------
void rebx_free_operator(struct rebx_operator* op) {
    if (!op) return;
    free(op->name);
    rebx_free_ap(&op->ap);
    free(op);
}

struct rebx_operator* rebx_create_operator(struct rebx_extras* const rebx, const char* name) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return nullptr; }
    struct rebx_operator* op = static_cast<struct rebx_operator*>(rebx_malloc(rebx, sizeof(*op)));
    if (!op) return nullptr;
    op->ap = nullptr; op->sim = rebx->sim; op->operator_type = REBX_OPERATOR_NONE;
    op->step_function = nullptr; op->free_memory = nullptr; op->name = nullptr;
    if (name) {
        op->name = static_cast<char*>(rebx_malloc(rebx, strlen(name) + 1));
        if (!op->name) { rebx_free_operator(op); return nullptr; }
        strcpy(op->name, name);
    }
    struct rebx_node* node = rebx_create_node(rebx);
    if (!node) { rebx_free_operator(op); return nullptr; }
    node->object = op;
    rebx_add_node(&rebx->allocated_operators, node);
    return op;
}

struct rebx_operator* rebx_load_operator(struct rebx_extras* const rebx, const char* name) {
    struct rebx_operator* op = rebx_create_operator(rebx, name);
    if (!op) return nullptr;
    if (strcmp(name, "integrate_force") == 0) { op->step_function = rebx_integrate_force; op->operator_type = REBX_OPERATOR_UPDATER; }
    else if (strcmp(name, "drift") == 0)      { op->step_function = rebx_drift_step;      op->operator_type = REBX_OPERATOR_UPDATER; }
    else if (strcmp(name, "kick") == 0)       { op->step_function = rebx_kick_step;       op->operator_type = REBX_OPERATOR_UPDATER; }
    else {
        char buf[300];
        snprintf(buf, sizeof buf, "REBOUNDx error: Operator '%s' not found in REBOUNDx library.\n", name);
        rebx_error(rebx, buf);
        rebx_remove_operator(rebx, op);
        return nullptr;
    }
    return op;
}

struct rebx_operator* rebx_get_operator(struct rebx_extras* const rebx, const char* const name) {
    for (struct rebx_node* n = rebx->allocated_operators; n; n = n->next) {
        struct rebx_operator* op = static_cast<struct rebx_operator*>(n->object);
        if (op->name && strcmp(op->name, name) == 0) return op;
    }
    return nullptr;
}

static int remove_steps_of(struct rebx_node** head, struct rebx_operator* op) {      // steps that use `op` (reference rebx_remove_operator)
    int removed = 0;
    for (struct rebx_node** link = head; *link;) {
        struct rebx_step* step = static_cast<struct rebx_step*>((*link)->object);
        if (step->operator_ == op) { struct rebx_node* dead = *link; *link = dead->next; free(step); free(dead); removed = 1; }
        else link = &(*link)->next;
    }
    return removed;
}

int rebx_remove_operator(struct rebx_extras* rebx, struct rebx_operator* op) {
    int removed = remove_steps_of(&rebx->pre_timestep_modifications, op);
    removed |= remove_steps_of(&rebx->post_timestep_modifications, op);
    if (rebx_remove_node(&rebx->allocated_operators, op)) { if (op->free_memory) op->free_memory(rebx, op); rebx_free_operator(op); removed = 1; }
    return removed;
}

int rebx_add_operator_step(struct rebx_extras* rebx, struct rebx_operator* op, const double dt_fraction, enum rebx_timing timing) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return 0; }
    if (op == nullptr) { rebx_error(rebx, "REBOUNDx error: Passed NULL pointer to rebx_add_operator_step.\n"); return 0; }
    if (op->step_function == nullptr) { rebx_error(rebx, "REBOUNDx error: Need to set step_function pointer on operator before adding to simulation. See custom effects example.\n"); return 0; }
    if (op->operator_type == REBX_OPERATOR_NONE) { rebx_error(rebx, "REBOUNDx error: Need to set operator_type field on operator before adding to simulation. See custom effects example.\n"); return 0; }
    if (timing != REBX_TIMING_PRE && timing != REBX_TIMING_POST) return 0;
    struct rebx_step* step = static_cast<struct rebx_step*>(rebx_malloc(rebx, sizeof(*step)));
    if (!step) return 0;
    step->operator_ = op; step->dt_fraction = dt_fraction;
    struct rebx_node* node = rebx_create_node(rebx);
    if (!node) { free(step); return 0; }
    node->object = step;
    struct reb_simulation* sim = rebx->sim;
    if (timing == REBX_TIMING_PRE) {
        rebx_add_node(&rebx->pre_timestep_modifications, node);
        if (sim->pre_timestep_modifications != nullptr && sim->pre_timestep_modifications != rebx_pre_timestep_modifications)
            reb_simulation_warning(sim, "REBOUNDx Warning: pre_timestep_modifications was set in the simulation and is being overwritten by REBOUNDx. To incorporate both, you can add your own custom effects through REBOUNDx.\n");
        sim->pre_timestep_modifications = rebx_pre_timestep_modifications;
    } else {
        rebx_add_node(&rebx->post_timestep_modifications, node);
        if (sim->post_timestep_modifications != nullptr && sim->post_timestep_modifications != rebx_post_timestep_modifications)
            reb_simulation_warning(sim, "REBOUNDx Warning: post_timestep_modifications was set in the simulation and is being overwritten by REBOUNDx. To incorporate both, you can add your own custom effects through REBOUNDx.\n");
        sim->post_timestep_modifications = rebx_post_timestep_modifications;
    }
    return 1;
}

int rebx_add_operator(struct rebx_extras* rebx, struct rebx_operator* op) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return 0; }
    if (op == nullptr) { rebx_error(rebx, "REBOUNDx error: Passed NULL pointer to rebx_add_operator.\n"); return 0; }
    struct reb_simulation* const sim = rebx->sim;
    if (op->operator_type == REBX_OPERATOR_RECORDER) return rebx_add_operator_step(rebx, op, 1., REBX_TIMING_POST);
    reb_simulation_warning(sim, "REBOUNDx Warning: Do not change the integrator after adding an operator (function assumed the current integrator when adding)");
    const char* in = sim->integrator.name ? sim->integrator.name : "";
    auto is = [&](const char* a) { return strcmp(in, a) == 0; };
    if (is("ias15") || is("bs")) return rebx_add_operator_step(rebx, op, 1., REBX_TIMING_POST);     // the step IAS15 / BS will take is not known in advance
    if (is("whfast") || is("saba") || is("leapfrog") || is("eos")) {                               // half a step before and after
        const int a = rebx_add_operator_step(rebx, op, 0.5, REBX_TIMING_PRE);
        const int b = rebx_add_operator_step(rebx, op, 0.5, REBX_TIMING_POST);
        return a && b;
    }
    if ((is("mercurius") || is("trace") || is("janus") || is("sei")) && op->operator_type == REBX_OPERATOR_UPDATER) {
        reb_simulation_error(sim, "REBOUNDx Error: Operators that affect particle trajectories are not supported with MERCURIUS, TRACE, SEI or Janus. Can only add forces.\n");
        return 0;
    }
    if (is("whfast512")) {
        reb_simulation_error(sim, "REBOUNDx Error: WHFast512 has been optimized for speed with options stripped out. This integrator will never be compatible with REBOUNDx.\n");
        return 0;
    }
    return 0;
}

static void run_steps(struct reb_simulation* sim, struct rebx_node* list, double dt) {
    if (sim->N_var != 0)
        reb_simulation_warning(sim, "REBOUNDx: Variational particles have been added to the simulation but are not implemented in REBOUNDx and will not be evolved self-consistently.");
    for (struct rebx_node* cur = list; cur; cur = cur->next) {
        struct rebx_step* step = static_cast<struct rebx_step*>(cur->object);
        struct rebx_operator* op = step->operator_;
        if (sim->integrator.name && strcmp(sim->integrator.name, "ias15") == 0 && op->operator_type == REBX_OPERATOR_UPDATER) {
            const struct reb_integrator_ias15_state* ias15 = static_cast<const struct reb_integrator_ias15_state*>(sim->integrator.state);
            if (ias15 && ias15->epsilon != 0)
                reb_simulation_warning(sim, "REBOUNDx: Operators that affect particle trajectories with adaptive timesteps can give spurious results. Use sim.ri_ias15.epsilon=0 for fixed timestep with IAS, or use a different integrator.");
        }
        op->step_function(sim, op, dt*step->dt_fraction);
    }
}

void rebx_pre_timestep_modifications(struct reb_simulation* sim) {
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    if (rebx) run_steps(sim, rebx->pre_timestep_modifications, sim->dt);
}

void rebx_post_timestep_modifications(struct reb_simulation* sim) {
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    if (rebx) run_steps(sim, rebx->post_timestep_modifications, sim->dt_last_done);
}

// reference src/steppers.c:109-130
void rebx_drift_step(struct reb_simulation* const sim, struct rebx_operator* const, const double dt) {
    struct reb_particle* const p = sim->particles;
    for (size_t i = 0; i < sim->N; i++) { p[i].x += dt*p[i].vx; p[i].y += dt*p[i].vy; p[i].z += dt*p[i].vz; }
}
void rebx_kick_step(struct reb_simulation* const sim, struct rebx_operator* const, const double dt) {
    reb_simulation_update_acceleration(sim);
    struct reb_particle* const p = sim->particles;
    for (size_t i = 0; i < sim->N; i++) { p[i].vx += dt*p[i].ax; p[i].vy += dt*p[i].ay; p[i].vz += dt*p[i].az; }
}

// Host schemes of integrate_force for a force that is not device-backed (a custom callback, a centre-of-mass frame):
// the reference's own sequence of force->update_accelerations calls on scratch copies (src/integrator_euler.c:33-41,
// integrator_rk2.c:38-66, integrator_rk4.c:40-90, integrator_implicit_midpoint.c:87-124).
static void host_integrate_force(struct reb_simulation* sim, struct rebx_force* force, double dt, int scheme) {
    const int N = (int)sim->N;
    struct reb_particle* const p = sim->particles;
    auto eval = [&](struct reb_particle* ps) { force->update_accelerations(sim, force, ps, N); };
    if (scheme == REBX_INTEGRATOR_EULER) {
        eval(p);
        for (int i = 0; i < N; i++) { p[i].vx += dt*p[i].ax; p[i].vy += dt*p[i].ay; p[i].vz += dt*p[i].az; }
    } else if (scheme == REBX_INTEGRATOR_RK2) {
        std::vector<struct reb_particle> k2(p, p + N);
        eval(p);
        const double a21 = 2.*dt/3.;
        for (int i = 0; i < N; i++) { k2[i].vx = p[i].vx + a21*p[i].ax; k2[i].vy = p[i].vy + a21*p[i].ay; k2[i].vz = p[i].vz + a21*p[i].az; }
        eval(k2.data());
        const double b1 = dt/4., b2 = 3.*dt/4.;
        for (int i = 0; i < N; i++) { p[i].vx += b1*p[i].ax + b2*k2[i].ax; p[i].vy += b1*p[i].ay + b2*k2[i].ay; p[i].vz += b1*p[i].az + b2*k2[i].az; }
    } else if (scheme == REBX_INTEGRATOR_RK4) {
        std::vector<struct reb_particle> k2(p, p + N), k3(p, p + N);
        const double dt2 = dt/2.;
        eval(p);
        for (int i = 0; i < N; i++) { k2[i].vx = p[i].vx + dt2*p[i].ax; k2[i].vy = p[i].vy + dt2*p[i].ay; k2[i].vz = p[i].vz + dt2*p[i].az; }
        eval(k2.data());
        for (int i = 0; i < N; i++) { k3[i].vx = p[i].vx + dt2*k2[i].ax; k3[i].vy = p[i].vy + dt2*k2[i].ay; k3[i].vz = p[i].vz + dt2*k2[i].az; }
        eval(k3.data());
        for (int i = 0; i < N; i++) {
            k2[i].vx = p[i].vx + dt*k3[i].ax; k2[i].vy = p[i].vy + dt*k3[i].ay; k2[i].vz = p[i].vz + dt*k3[i].az;
            k3[i].ax += k2[i].ax; k3[i].ay += k2[i].ay; k3[i].az += k2[i].az;         // a2 + a3, before k2 is reused for stage 4
            k2[i].ax = 0.; k2[i].ay = 0.; k2[i].az = 0.;
        }
        eval(k2.data());
        const double w = dt/6.;
        for (int i = 0; i < N; i++) {
            p[i].vx += w*(p[i].ax + k2[i].ax + 2.*k3[i].ax); p[i].vy += w*(p[i].ay + k2[i].ay + 2.*k3[i].ay); p[i].vz += w*(p[i].az + k2[i].az + 2.*k3[i].az);
        }
    } else if (scheme == REBX_INTEGRATOR_IMPLICIT_MIDPOINT) {
        std::vector<struct reb_particle> fin(p, p + N), prev(N), avg(p, p + N);
        int n;
        for (n = 0; n < 10; n++) {
            prev = fin;
            eval(avg.data());
            double d2 = 0., t2 = 0.;
            for (int i = 0; i < N; i++) {
                fin[i].vx = p[i].vx + dt*avg[i].ax; fin[i].vy = p[i].vy + dt*avg[i].ay; fin[i].vz = p[i].vz + dt*avg[i].az;
                const double dx = fin[i].vx - prev[i].vx, dy = fin[i].vy - prev[i].vy, dz = fin[i].vz - prev[i].vz;
                d2 += dx*dx + dy*dy + dz*dz;
                t2 += fin[i].vx*fin[i].vx + fin[i].vy*fin[i].vy + fin[i].vz*fin[i].vz;
            }
            if (d2/t2 < 2.220446049250313e-16*2.220446049250313e-16) break;
            for (int i = 0; i < N; i++) {
                avg[i].x = 0.5*(p[i].x + fin[i].x); avg[i].y = 0.5*(p[i].y + fin[i].y); avg[i].z = 0.5*(p[i].z + fin[i].z);
                avg[i].vx = 0.5*(p[i].vx + fin[i].vx); avg[i].vy = 0.5*(p[i].vy + fin[i].vy); avg[i].vz = 0.5*(p[i].vz + fin[i].vz);
                avg[i].ax = 0.; avg[i].ay = 0.; avg[i].az = 0.;
                avg[i].m = 0.5*(p[i].m + fin[i].m);
            }
        }
        if (n == 10)
            reb_simulation_warning(sim, "REBOUNDx: 10 iterations in integrator_implicit_midpoint.c failed to converge. This is typically because the perturbation is too strong for the current implementation.");
        for (int i = 0; i < N; i++) { p[i].vx = fin[i].vx; p[i].vy = fin[i].vy; p[i].vz = fin[i].vz; }
    }
}

// reference src/integrate_force.c:34-78.  Device-backed force in PARTICLE coordinates about particle 0: the particles go
// to the device once, all stages of the scheme run there (rebx_b200_integrate_force), velocities and the stage-1
// accelerations come back.  Any other force: the host schemes above, calling force->update_accelerations.
void rebx_integrate_force(struct reb_simulation* const sim, struct rebx_operator* const op, const double dt) {
    using namespace rebxh;
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    struct rebx_force* force = static_cast<struct rebx_force*>(find_value(op->ap, "force"));
    if (force == nullptr) {
        reb_simulation_error(sim, "REBOUNDx Error: Force parameter not set in rebx_integrate operator. See examples for how to add as a parameter.\n");
        return;
    }
    int scheme = REBX_INTEGRATOR_EULER;
    if (const int* ip = static_cast<const int*>(find_value(op->ap, "integrator"))) scheme = *ip;
    rebx_reset_accelerations(sim->particles, (int)sim->N);
    if (scheme < 0 || scheme > 3) return;                 // REBX_INTEGRATOR_NONE and unknown values: nothing to do (reference :73-79)
    const int kind = kind_of(force);
    bool on_device = kind != REBX_B200_NONE && kind != KIND_GR && sim->N >= 2;
    if (on_device) {
        const int* cptr = static_cast<const int*>(find_value(force->ap, "coordinates"));
        const bool com_force = kind == REBX_B200_MODIFY_ORBITS || kind == REBX_B200_GAS_DAMPING || kind == REBX_B200_TYPE_I || kind == REBX_B200_EXP_MIGRATION;
        if (com_force && (!cptr || *cptr != REBX_COORDINATES_PARTICLE)) on_device = false;      // centre-of-mass frames: host schemes around the device sweeps
    }
    if (on_device) {
        // a session needs every effect entry of the force to refer to particle 0 (several radiation sources or oblate
        // bodies, a primary other than particle 0: host schemes around the device sweeps instead)
        Side* side = side_of(rebx);
        std::string why;
        DeviceContext* ctx = ensure_ctx(side, why);
        std::vector<Entry> probe;
        if (ctx) entries_for_force(sim, side, ctx, force, kind, sim->particles, sim->N, probe, ctx->tables);
        if (!ctx || probe.empty() || probe.size() > REBX_B200_MAX_EFFECTS) on_device = false;
        for (const Entry& e : probe) if (e.frame != 0 || e.body_index != 0) on_device = false;
        if (ctx && probe.empty()) return;          // the force reported why it does nothing (or has nothing to act on)
    }
    if (on_device) {
        struct rebx_force* one = force;
        rebx_b200_session* s = session_open(sim, rebx, &one, 1);
        if (s) {
            // only velocities (and the stage-1 accelerations, as the reference leaves them in sim->particles) change
            int rc = rebx_b200_session_integrate_force(s, dt, scheme);
            if (rc == 0) {
                const rebx_b200_state& st = s->P.st;
                rebx_b200_aos_layout L = kLayout;
                rebx_b200_state v = st;
                L.off_a = kLayout.off_v; v.ax = const_cast<double*>(st.vx); v.ay = const_cast<double*>(st.vy); v.az = const_cast<double*>(st.vz);
                rc = rebx_b200_pack_acc_aos(s->aos.p, &L, &v, s->stream, &s->launches_i);
                if (rc == 0) rc = rebx_b200_pack_acc_aos(s->aos.p, &kLayout, &st, s->stream, &s->launches_i);
                if (rc == 0) rc = (int)copy_page_split(sim->particles, s->aos.p, s->N*sizeof(struct reb_particle), cudaMemcpyDeviceToHost, s->stream);
                if (rc == 0) rc = (int)cudaStreamSynchronize(s->stream);
                if (rc != 0) session_fail(s, "copying the velocities to the host", rc);
                s->ctx->stats.d2h_bytes += s->N*sizeof(struct reb_particle);
                session_count(s);
            }
            rebx_b200_session_free(s);
            return;
        }
        return;      // the session reported why it could not be opened; like the reference on an error, leave the state alone
    }
    host_integrate_force(sim, force, dt, scheme);
}

}  // extern "C"
