// extern "C" surface of libtmb200.so (declared in include/tmb200.h).  Thin: argument marshalling, exception ->
// status code + message, handle bookkeeping.  All compute goes through the classes in potential.hpp.
#include "../../include/tmb200.h"
#include "fixed_point.cuh"
#include "exchange.hpp"
#include "potential.hpp"

#include <cstring>

using namespace tmb;

static thread_local std::string g_last_error;

template <typename F> static int guarded(F &&f) {
    try {
        f();
        return TMB_OK;
    } catch (const InvalidHardware &e) {
        g_last_error = e.what();
        return TMB_INVALID_HARDWARE;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return TMB_ERROR;
    } catch (...) {
        g_last_error = "unknown error";
        return TMB_ERROR;
    }
}

typedef std::shared_ptr<Potential> PotPtr;
typedef std::shared_ptr<BoundPotential> BpPtr;
typedef std::shared_ptr<Integrator> IntgPtr;

static PotPtr &as_pot(tmb_potential h) {
    if (h == nullptr) {
        throw std::runtime_error("null potential handle");
    }
    return *static_cast<PotPtr *>(h);
}
static BpPtr &as_bp(tmb_bound_potential h) {
    if (h == nullptr) {
        throw std::runtime_error("null bound potential handle");
    }
    return *static_cast<BpPtr *>(h);
}
static IntgPtr &as_intg(tmb_integrator h) {
    if (h == nullptr) {
        throw std::runtime_error("null integrator handle");
    }
    return *static_cast<IntgPtr *>(h);
}
static LangevinIntegrator &as_langevin(tmb_integrator h) {
    auto p = std::dynamic_pointer_cast<LangevinIntegrator>(as_intg(h));
    if (p == nullptr) {
        throw std::runtime_error("integrator must be LangevinIntegrator.");
    }
    return *p;
}
static Context &as_ctx(tmb_context h) {
    if (h == nullptr) {
        throw std::runtime_error("null context handle");
    }
    return *static_cast<Context *>(h);
}

typedef std::shared_ptr<Mover> MoverPtr;
static MoverPtr &as_mover(tmb_mover h) {
    if (h == nullptr) {
        throw std::runtime_error("null mover handle");
    }
    return *static_cast<MoverPtr *>(h);
}
static std::shared_ptr<MonteCarloBarostat<float>> as_barostat(tmb_mover h) {
    auto b = std::dynamic_pointer_cast<MonteCarloBarostat<float>>(as_mover(h));
    if (!b) {
        throw std::runtime_error("mover is not a MonteCarloBarostat");
    }
    return b;
}

struct NeighborlistHandle {
    int precision;
    std::unique_ptr<Neighborlist<float>> f32;
    std::unique_ptr<Neighborlist<double>> f64;
    std::vector<std::vector<int>> last; // result of the last build, served by fetch
};
static NeighborlistHandle &as_nb(tmb_neighborlist h) {
    if (h == nullptr) {
        throw std::runtime_error("null neighborlist handle");
    }
    return *static_cast<NeighborlistHandle *>(h);
}

static void check_precision(int precision) {
    if (precision != TMB_F32 && precision != TMB_F64) {
        throw std::runtime_error("precision must be 32 or 64");
    }
}

static i128 *as_i128(tmb_i128 *p) {
    static_assert(sizeof(tmb_i128) == sizeof(i128), "tmb_i128 must alias __int128");
    return reinterpret_cast<i128 *>(p);
}

template <template <typename> class PotT>
static int create_bonded(int precision, const int32_t *idxs, int n_values, tmb_potential *out) {
    return guarded([&] {
        check_precision(precision);
        std::vector<int> v(idxs, idxs + (n_values > 0 ? n_values : 0));
        PotPtr p;
        if (precision == TMB_F32) {
            p = std::make_shared<PotT<float>>(v);
        } else {
            p = std::make_shared<PotT<double>>(v);
        }
        *out = new PotPtr(p);
    });
}

template <template <typename> class PotT>
static int create_restraint(
    int precision, const int32_t *idxs, int n_values, const int32_t *signs, int n_signs, double beta, double cutoff,
    tmb_potential *out) {
    return guarded([&] {
        check_precision(precision);
        std::vector<int> v(idxs, idxs + (n_values > 0 ? n_values : 0));
        std::vector<int> sg(signs, signs + (n_signs > 0 ? n_signs : 0));
        PotPtr p;
        if (precision == TMB_F32) {
            p = std::make_shared<PotT<float>>(v, sg, beta, cutoff);
        } else {
            p = std::make_shared<PotT<double>>(v, sg, beta, cutoff);
        }
        *out = new PotPtr(p);
    });
}
extern "C" {

const char *tmb_last_error(void) { return g_last_error.c_str(); }
int tmb_version(void) { return 1; }
uint64_t tmb_fixed_exponent(void) { return FIXED_EXPONENT; }
long long tmb_kernel_launch_count(void) { return g_kernel_launches.load(); }

int tmb_cuda_device_reset(void) {
    return guarded([&] { TMB_CUDA(cudaDeviceReset()); });
}
int tmb_set_stream(void *cuda_stream) {
    return guarded([&] { set_main_stream(static_cast<cudaStream_t>(cuda_stream)); });
}
int tmb_device_synchronize(void) {
    return guarded([&] { TMB_CUDA(cudaDeviceSynchronize()); });
}

// ---- constructors ---------------------------------------------------------------------------------------------
int tmb_harmonic_bond_create(int precision, const int32_t *bond_idxs, int n_values, tmb_potential *out) {
    return create_bonded<HarmonicBond>(precision, bond_idxs, n_values, out);
}
int tmb_harmonic_angle_create(int precision, const int32_t *angle_idxs, int n_values, tmb_potential *out) {
    return create_bonded<HarmonicAngle>(precision, angle_idxs, n_values, out);
}
int tmb_periodic_torsion_create(int precision, const int32_t *torsion_idxs, int n_values, tmb_potential *out) {
    return create_bonded<PeriodicTorsion>(precision, torsion_idxs, n_values, out);
}

int tmb_flat_bottom_bond_create(int precision, const int32_t *bond_idxs, int n_values, tmb_potential *out) {
    return create_restraint<FlatBottomBond>(precision, bond_idxs, n_values, nullptr, 0, 0.0, 0.0, out);
}
int tmb_log_flat_bottom_bond_create(int precision, const int32_t *bond_idxs, int n_values, double beta, tmb_potential *out) {
    return create_restraint<LogFlatBottomBond>(precision, bond_idxs, n_values, nullptr, 0, beta, 0.0, out);
}
int tmb_centroid_restraint_create(
    int precision, const int32_t *group_a_idxs, int n_a, const int32_t *group_b_idxs, int n_b, double kb, double b0,
    tmb_potential *out) {
    return guarded([&] {
        const std::vector<int> ga(group_a_idxs, group_a_idxs + n_a), gb(group_b_idxs, group_b_idxs + n_b);
        if (precision == TMB_F32) {
            *out = new PotPtr(std::make_shared<CentroidRestraint<float>>(ga, gb, kb, b0));
        } else {
            *out = new PotPtr(std::make_shared<CentroidRestraint<double>>(ga, gb, kb, b0));
        }
    });
}
int tmb_chiral_atom_restraint_create(int precision, const int32_t *idxs, int n_values, tmb_potential *out) {
    return create_restraint<ChiralAtomRestraint>(precision, idxs, n_values, nullptr, 0, 0.0, 0.0, out);
}
int tmb_chiral_bond_restraint_create(
    int precision, const int32_t *idxs, int n_values, const int32_t *signs, int n_signs, tmb_potential *out) {
    return create_restraint<ChiralBondRestraint>(precision, idxs, n_values, signs, n_signs, 0.0, 0.0, out);
}
int tmb_nonbonded_pair_list_precomputed_create(
    int precision, const int32_t *pair_idxs, int n_values, double beta, double cutoff, tmb_potential *out) {
    return create_restraint<NonbondedPairListPrecomputed>(precision, pair_idxs, n_values, nullptr, 0, beta, cutoff, out);
}

int tmb_nonbonded_all_pairs_create(
    int precision, int num_atoms, double beta, double cutoff, const int32_t *atom_idxs, int n_atom_idxs,
    int disable_hilbert_sort, double nblist_padding, tmb_potential *out) {
    return guarded([&] {
        check_precision(precision);
        std::optional<std::set<int>> idxs;
        if (n_atom_idxs >= 0) {
            // the reference converts the array to a std::set (dedup + ascending) before validating, wrap_kernels.cpp:1460
            idxs = std::set<int>(atom_idxs, atom_idxs + n_atom_idxs);
            std::vector<int> raw(atom_idxs, atom_idxs + n_atom_idxs);
            verify_atom_idxs(num_atoms, raw);
        }
        PotPtr p;
        if (precision == TMB_F32) {
            p = std::make_shared<NonbondedAllPairs<float>>(num_atoms, beta, cutoff, idxs, disable_hilbert_sort != 0, nblist_padding);
        } else {
            p = std::make_shared<NonbondedAllPairs<double>>(num_atoms, beta, cutoff, idxs, disable_hilbert_sort != 0, nblist_padding);
        }
        *out = new PotPtr(p);
    });
}

int tmb_nonbonded_all_pairs_set_atom_idxs(tmb_potential pot, const int32_t *atom_idxs, int n) {
    return guarded([&] {
        std::vector<int> v(atom_idxs, atom_idxs + (n > 0 ? n : 0));
        if (auto a = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(as_pot(pot))) {
            a->set_atom_idxs(v);
        } else if (auto b = std::dynamic_pointer_cast<NonbondedAllPairs<double>>(as_pot(pot))) {
            b->set_atom_idxs(v);
        } else {
            throw std::runtime_error("not a NonbondedAllPairs potential");
        }
    });
}

int tmb_nonbonded_all_pairs_get_num_atom_idxs(tmb_potential pot, int *out) {
    return guarded([&] {
        if (auto a = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(as_pot(pot))) {
            *out = a->get_num_atom_idxs();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedAllPairs<double>>(as_pot(pot))) {
            *out = b->get_num_atom_idxs();
        } else {
            throw std::runtime_error("not a NonbondedAllPairs potential");
        }
    });
}

int tmb_nonbonded_all_pairs_get_atom_idxs(tmb_potential pot, int32_t *out) {
    return guarded([&] {
        std::vector<int> v;
        if (auto a = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(as_pot(pot))) {
            v = a->get_atom_idxs();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedAllPairs<double>>(as_pot(pot))) {
            v = b->get_atom_idxs();
        } else {
            throw std::runtime_error("not a NonbondedAllPairs potential");
        }
        std::copy(v.begin(), v.end(), out);
    });
}

int tmb_nonbonded_interaction_group_create(
    int precision, int num_atoms, const int32_t *row_atom_idxs, int n_row, double beta, double cutoff,
    const int32_t *col_atom_idxs, int n_col, int disable_hilbert_sort, double nblist_padding, tmb_potential *out) {
    return guarded([&] {
        check_precision(precision);
        std::vector<int> rows(row_atom_idxs, row_atom_idxs + (n_row > 0 ? n_row : 0));
        std::vector<int> cols;
        if (n_col >= 0) {
            cols.assign(col_atom_idxs, col_atom_idxs + n_col);
        } else {
            // None: columns are the complement of the rows (wrap_kernels.cpp:1496-1502)
            std::set<int> rs(rows.begin(), rows.end());
            for (int i = 0; i < num_atoms; i++) {
                if (!rs.count(i)) {
                    cols.push_back(i);
                }
            }
        }
        PotPtr p;
        if (precision == TMB_F32) {
            p = std::make_shared<NonbondedInteractionGroup<float>>(
                num_atoms, rows, cols, beta, cutoff, disable_hilbert_sort != 0, nblist_padding);
        } else {
            p = std::make_shared<NonbondedInteractionGroup<double>>(
                num_atoms, rows, cols, beta, cutoff, disable_hilbert_sort != 0, nblist_padding);
        }
        *out = new PotPtr(p);
    });
}

int tmb_nonbonded_interaction_group_set_atom_idxs(
    tmb_potential pot, const int32_t *row_atom_idxs, int n_row, const int32_t *col_atom_idxs, int n_col) {
    return guarded([&] {
        std::vector<int> rows(row_atom_idxs, row_atom_idxs + (n_row > 0 ? n_row : 0));
        std::vector<int> cols(col_atom_idxs, col_atom_idxs + (n_col > 0 ? n_col : 0));
        if (auto a = std::dynamic_pointer_cast<NonbondedInteractionGroup<float>>(as_pot(pot))) {
            a->set_atom_idxs(rows, cols);
        } else if (auto b = std::dynamic_pointer_cast<NonbondedInteractionGroup<double>>(as_pot(pot))) {
            b->set_atom_idxs(rows, cols);
        } else {
            throw std::runtime_error("not a NonbondedInteractionGroup potential");
        }
    });
}

int tmb_nonbonded_pair_list_create(
    int precision, int negated, const int32_t *pair_idxs, int n_pair_values, const double *scales, int n_scale_values,
    double beta, double cutoff, tmb_potential *out) {
    return guarded([&] {
        check_precision(precision);
        std::vector<int> pairs(pair_idxs, pair_idxs + (n_pair_values > 0 ? n_pair_values : 0));
        std::vector<double> sc(scales, scales + (n_scale_values > 0 ? n_scale_values : 0));
        PotPtr p;
        if (precision == TMB_F32) {
            if (negated) {
                p = std::make_shared<NonbondedPairList<float, true>>(pairs, sc, beta, cutoff);
            } else {
                p = std::make_shared<NonbondedPairList<float, false>>(pairs, sc, beta, cutoff);
            }
        } else {
            if (negated) {
                p = std::make_shared<NonbondedPairList<double, true>>(pairs, sc, beta, cutoff);
            } else {
                p = std::make_shared<NonbondedPairList<double, false>>(pairs, sc, beta, cutoff);
            }
        }
        *out = new PotPtr(p);
    });
}

int tmb_summed_potential_create(
    const tmb_potential *potentials, int n_potentials, const int32_t *params_sizes, int n_sizes, int parallel,
    tmb_potential *out) {
    return guarded([&] {
        std::vector<PotPtr> pots;
        for (int i = 0; i < n_potentials; i++) {
            pots.push_back(as_pot(potentials[i]));
        }
        std::vector<int> sizes(params_sizes, params_sizes + (n_sizes > 0 ? n_sizes : 0));
        *out = new PotPtr(std::make_shared<SummedPotential>(pots, sizes, parallel != 0));
    });
}

int tmb_fanout_summed_potential_create(const tmb_potential *potentials, int n_potentials, int parallel, tmb_potential *out) {
    return guarded([&] {
        std::vector<PotPtr> pots;
        for (int i = 0; i < n_potentials; i++) {
            pots.push_back(as_pot(potentials[i]));
        }
        *out = new PotPtr(std::make_shared<FanoutSummedPotential>(pots, parallel != 0));
    });
}

int tmb_potential_destroy(tmb_potential pot) {
    return guarded([&] { delete static_cast<PotPtr *>(pot); });
}

// ---- evaluation -----------------------------------------------------------------------------------------------
int tmb_potential_execute(
    tmb_potential pot, int N, int P, const double *coords, const double *params, const double *box, uint64_t *du_dx,
    uint64_t *du_dp, tmb_i128 *u) {
    return guarded([&] {
        as_pot(pot)->execute_host(N, P, coords, params, box, reinterpret_cast<u64 *>(du_dx), reinterpret_cast<u64 *>(du_dp), as_i128(u));
    });
}

int tmb_potential_execute_batch(
    tmb_potential pot, int coord_batches, int N, int param_batches, int P, const double *coords, const double *params,
    const double *boxes, uint64_t *du_dx, uint64_t *du_dp, tmb_i128 *u) {
    return guarded([&] {
        as_pot(pot)->execute_batch_host(
            coord_batches, N, param_batches, P, coords, params, boxes, reinterpret_cast<u64 *>(du_dx),
            reinterpret_cast<u64 *>(du_dp), as_i128(u));
    });
}

int tmb_potential_execute_batch_sparse(
    tmb_potential pot, int coords_size, int N, int params_size, int P, int batch_size, const uint32_t *coords_batch_idxs,
    const uint32_t *params_batch_idxs, const double *coords, const double *params, const double *boxes, uint64_t *du_dx,
    uint64_t *du_dp, tmb_i128 *u) {
    return guarded([&] {
        as_pot(pot)->execute_batch_sparse_host(
            coords_size, N, params_size, P, batch_size, coords_batch_idxs, params_batch_idxs, coords, params, boxes,
            reinterpret_cast<u64 *>(du_dx), reinterpret_cast<u64 *>(du_dp), as_i128(u));
    });
}

int tmb_potential_du_dp_fixed_to_float(tmb_potential pot, int N, int P, const uint64_t *du_dp, double *out) {
    return guarded([&] { as_pot(pot)->du_dp_fixed_to_float(N, P, reinterpret_cast<const u64 *>(du_dp), out); });
}

int tmb_potential_execute_device(
    tmb_potential pot, int N, int P, const double *d_coords, const double *d_params, const double *d_box,
    uint64_t *d_du_dx, uint64_t *d_du_dp, tmb_i128 *d_u, void *cuda_stream) {
    return guarded([&] {
        as_pot(pot)->execute_device(
            N, P, d_coords, d_params, d_box, reinterpret_cast<u64 *>(d_du_dx), reinterpret_cast<u64 *>(d_du_dp),
            as_i128(d_u), static_cast<cudaStream_t>(cuda_stream));
    });
}

int tmb_nonbonded_num_tiles(tmb_potential pot, unsigned int *out) {
    return guarded([&] {
        if (auto a = std::dynamic_pointer_cast<NonbondedTiled<float>>(as_pot(pot))) {
            *out = a->num_tiles();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedTiled<double>>(as_pot(pot))) {
            *out = b->num_tiles();
        } else {
            throw std::runtime_error("not a tile-list nonbonded potential");
        }
    });
}

int tmb_nonbonded_tile_capacity(tmb_potential pot, unsigned long long *capacity, unsigned long long *worst_case) {
    return guarded([&] {
        if (auto a = std::dynamic_pointer_cast<NonbondedTiled<float>>(as_pot(pot))) {
            *capacity = a->tile_capacity();
            *worst_case = a->tile_worst_case();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedTiled<double>>(as_pot(pot))) {
            *capacity = b->tile_capacity();
            *worst_case = b->tile_worst_case();
        } else {
            throw std::runtime_error("not a tile-list nonbonded potential");
        }
    });
}

int tmb_nonbonded_num_rebuilds(tmb_potential pot, unsigned int *out) {
    return guarded([&] {
        if (auto a = std::dynamic_pointer_cast<NonbondedTiled<float>>(as_pot(pot))) {
            *out = a->num_rebuilds();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedTiled<double>>(as_pot(pot))) {
            *out = b->num_rebuilds();
        } else {
            throw std::runtime_error("not a tile-list nonbonded potential");
        }
    });
}

int tmb_nonbonded_set_kernel_timing(tmb_potential pot, int on) {
    return guarded([&] {
        if (auto a = std::dynamic_pointer_cast<NonbondedTiled<float>>(as_pot(pot))) {
            a->set_kernel_timing(on != 0);
        } else if (auto b = std::dynamic_pointer_cast<NonbondedTiled<double>>(as_pot(pot))) {
            b->set_kernel_timing(on != 0);
        } else {
            throw std::runtime_error("not a tile-list nonbonded potential");
        }
    });
}

int tmb_nonbonded_drain_kernel_times(tmb_potential pot, float *out_ms, int capacity, int *n_out) {
    return guarded([&] {
        std::vector<float> t;
        if (auto a = std::dynamic_pointer_cast<NonbondedTiled<float>>(as_pot(pot))) {
            t = a->drain_kernel_times();
        } else if (auto b = std::dynamic_pointer_cast<NonbondedTiled<double>>(as_pot(pot))) {
            t = b->drain_kernel_times();
        } else {
            throw std::runtime_error("not a tile-list nonbonded potential");
        }
        const int n = std::min<int>(capacity, static_cast<int>(t.size()));
        std::copy(t.begin(), t.begin() + n, out_ms);
        *n_out = n;
    });
}

// ---- BoundPotential ---------------------------------------------------------------------------------------------
int tmb_bound_potential_create(tmb_potential pot, const double *params, int n_params, tmb_bound_potential *out) {
    return guarded([&] {
        std::vector<double> p(params, params + (n_params > 0 ? n_params : 0));
        *out = new BpPtr(std::make_shared<BoundPotential>(as_pot(pot), p));
    });
}
int tmb_bound_potential_destroy(tmb_bound_potential bp) {
    return guarded([&] { delete static_cast<BpPtr *>(bp); });
}
int tmb_bound_potential_set_params(tmb_bound_potential bp, const double *params, int n_params) {
    return guarded([&] {
        std::vector<double> p(params, params + (n_params > 0 ? n_params : 0));
        as_bp(bp)->set_params(p);
    });
}
int tmb_bound_potential_size(tmb_bound_potential bp, int *out) {
    return guarded([&] { *out = as_bp(bp)->size; });
}
int tmb_bound_potential_execute(
    tmb_bound_potential bp, int N, const double *coords, const double *box, uint64_t *du_dx, tmb_i128 *u) {
    return guarded([&] { as_bp(bp)->execute_host(N, coords, box, reinterpret_cast<u64 *>(du_dx), as_i128(u)); });
}
int tmb_bound_potential_execute_batch(
    tmb_bound_potential bp, int coord_batches, int N, const double *coords, const double *boxes, uint64_t *du_dx,
    tmb_i128 *u) {
    return guarded([&] {
        as_bp(bp)->execute_batch_host(coord_batches, N, coords, boxes, reinterpret_cast<u64 *>(du_dx), as_i128(u));
    });
}
int tmb_bound_potential_set_params_device(tmb_bound_potential bp, const double *d_params, int n_params, void *cuda_stream) {
    return guarded([&] { as_bp(bp)->set_params_device(n_params, d_params, static_cast<cudaStream_t>(cuda_stream)); });
}
int tmb_bound_potential_execute_device(
    tmb_bound_potential bp, int N, const double *d_coords, const double *d_box, uint64_t *d_du_dx, tmb_i128 *d_u,
    void *cuda_stream) {
    return guarded([&] {
        as_bp(bp)->execute_device(
            N, d_coords, d_box, reinterpret_cast<u64 *>(d_du_dx), nullptr, as_i128(d_u), static_cast<cudaStream_t>(cuda_stream));
    });
}

// ---- integrator / context -----------------------------------------------------------------------------------------
int tmb_langevin_integrator_create(
    const double *masses, int N, double temperature, double dt, double friction, int seed, tmb_integrator *out) {
    return guarded([&] { *out = new IntgPtr(std::make_shared<LangevinIntegrator>(N, masses, temperature, dt, friction, seed)); });
}
int tmb_langevin_integrator_destroy(tmb_integrator intg) {
    return guarded([&] { delete static_cast<IntgPtr *>(intg); });
}
int tmb_langevin_integrator_set_noise(tmb_integrator intg, const float *noise) {
    return guarded([&] { as_langevin(intg).set_external_noise(noise); });
}
int tmb_langevin_integrator_set_step(tmb_integrator intg, unsigned long long step) {
    return guarded([&] { as_langevin(intg).set_step(static_cast<long long>(step)); });
}
int tmb_langevin_integrator_get_step(tmb_integrator intg, unsigned long long *step) {
    return guarded([&] { *step = static_cast<unsigned long long>(as_langevin(intg).step_count()); });
}

int tmb_rmsd_align(const double *x1, const double *x2, int N, double *x2_aligned) {
    return guarded([&] { rmsd_align_host(N, x1, x2, x2_aligned); });
}
int tmb_velocity_verlet_integrator_create(double dt, const double *cbs, int N, tmb_integrator *out) {
    return guarded([&] { *out = new IntgPtr(std::make_shared<VelocityVerletIntegrator>(N, dt, cbs)); });
}
int tmb_integrator_destroy(tmb_integrator intg) {
    return guarded([&] { delete static_cast<IntgPtr *>(intg); });
}

int tmb_context_create(
    const double *x0, const double *v0, const double *box, int N, tmb_integrator intg, const tmb_bound_potential *bps,
    int n_bps, tmb_context *out) {
    return guarded([&] {
        std::vector<BpPtr> v;
        for (int i = 0; i < n_bps; i++) {
            v.push_back(as_bp(bps[i]));
        }
        *out = new Context(N, x0, v0, box, as_intg(intg), v);
    });
}
int tmb_context_create_with_movers(
    const double *x0, const double *v0, const double *box, int N, tmb_integrator intg, const tmb_bound_potential *bps,
    int n_bps, const tmb_mover *movers, int n_movers, tmb_context *out) {
    return guarded([&] {
        std::vector<BpPtr> v;
        for (int i = 0; i < n_bps; i++) {
            v.push_back(as_bp(bps[i]));
        }
        std::vector<MoverPtr> mv;
        for (int i = 0; i < n_movers; i++) {
            mv.push_back(as_mover(movers[i]));
        }
        *out = new Context(N, x0, v0, box, as_intg(intg), v, mv);
    });
}
int tmb_barostat_create(
    int N, double pressure, double temperature, const int *group_atoms, const int *group_offsets, int n_groups,
    int interval, const tmb_bound_potential *bps, int n_bps, int seed, int adaptive_scaling_enabled,
    double initial_volume_scale_factor, tmb_mover *out) {
    return guarded([&] {
        std::vector<std::vector<int>> groups(n_groups);
        for (int g = 0; g < n_groups; g++) {
            groups[g].assign(group_atoms + group_offsets[g], group_atoms + group_offsets[g + 1]);
        }
        std::vector<BpPtr> v;
        for (int i = 0; i < n_bps; i++) {
            v.push_back(as_bp(bps[i]));
        }
        *out = new MoverPtr(std::make_shared<MonteCarloBarostat<float>>(
            N, pressure, temperature, groups, interval, v, seed, adaptive_scaling_enabled != 0, initial_volume_scale_factor));
    });
}
// ---- water exchange (biased deletion) -----------------------------------------------------------------------------
extern "C++" {
static std::vector<std::vector<int>> unflatten_mols(const int *atoms, const int *offsets, int n_mols) {
    std::vector<std::vector<int>> mols(n_mols > 0 ? n_mols : 0);
    for (int m = 0; m < n_mols; m++) {
        mols[m].assign(atoms + offsets[m], atoms + offsets[m + 1]);
    }
    return mols;
}

// f32 and f64 movers behind one set of entry points
template <typename F32, typename F64> static auto with_bd(tmb_mover h, F32 &&f32, F64 &&f64) {
    MoverPtr &m = as_mover(h);
    if (auto a = std::dynamic_pointer_cast<BDExchangeMove<float>>(m)) {
        return f32(*a);
    }
    if (auto b = std::dynamic_pointer_cast<BDExchangeMove<double>>(m)) {
        return f64(*b);
    }
    throw std::runtime_error("mover is not a BDExchangeMove");
}
#define WITH_BD(h, expr) with_bd((h), [&](BDExchangeMove<float> &mv) { return expr; }, [&](BDExchangeMove<double> &mv) { return expr; })

template <typename Real> static void widen(const std::vector<Real> &v, double *out) {
    for (size_t i = 0; i < v.size(); i++) {
        out[i] = static_cast<double>(v[i]);
    }
}

} // extern "C++"
int tmb_bd_exchange_move_create(
    int precision, int N, const int *mol_atoms, const int *mol_offsets, int n_mols, const double *params, int n_params,
    double temperature, double nb_beta, double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size,
    tmb_mover *out) {
    return guarded([&] {
        check_precision(precision);
        // argument checks of the reference's binding, in its order (wrap_kernels.cpp:1749-1769)
        if (num_proposals_per_move <= 0) {
            throw std::runtime_error("proposals per move must be greater than 0");
        }
        if (n_params != N * P_PER_ATOM) {
            throw std::runtime_error("Number of parameters must match N");
        }
        if (n_mols <= 0) {
            throw std::runtime_error("must provide at least one molecule");
        }
        if (interval <= 0) {
            throw std::runtime_error("must provide interval greater than 0");
        }
        if (batch_size <= 0) {
            throw std::runtime_error("must provide batch size greater than 0");
        }
        if (batch_size > num_proposals_per_move) {
            throw std::runtime_error("number of proposals per move must be greater than batch size");
        }
        auto mols = unflatten_mols(mol_atoms, mol_offsets, n_mols);
        std::vector<double> p(params, params + n_params);
        MoverPtr mv;
        if (precision == TMB_F32) {
            mv = std::make_shared<BDExchangeMove<float>>(N, mols, p, temperature, nb_beta, cutoff, seed, num_proposals_per_move, interval, batch_size);
        } else {
            mv = std::make_shared<BDExchangeMove<double>>(N, mols, p, temperature, nb_beta, cutoff, seed, num_proposals_per_move, interval, batch_size);
        }
        *out = new MoverPtr(mv);
    });
}
int tmb_tibd_exchange_move_create(
    int precision, int N, const int *ligand_idxs, int n_ligand, const int *mol_atoms, const int *mol_offsets, int n_mols,
    const double *params, int n_params, double temperature, double nb_beta, double cutoff, double radius, int seed,
    int num_proposals_per_move, int interval, int batch_size, tmb_mover *out) {
    return guarded([&] {
        check_precision(precision);
        // argument checks of the reference's binding, in its order (wrap_kernels.cpp:1922-1947)
        if (num_proposals_per_move <= 0) {
            throw std::runtime_error("proposals per move must be greater than 0");
        }
        if (n_params != N * P_PER_ATOM) {
            throw std::runtime_error("Number of parameters must match N");
        }
        if (n_ligand <= 0) {
            throw std::runtime_error("must provide at least one atom for the ligand indices");
        }
        if (n_mols <= 0) {
            throw std::runtime_error("must provide at least one molecule");
        }
        if (interval <= 0) {
            throw std::runtime_error("must provide interval greater than 0");
        }
        if (batch_size <= 0) {
            throw std::runtime_error("must provide batch size greater than 0");
        }
        if (batch_size > num_proposals_per_move) {
            throw std::runtime_error("number of proposals per move must be greater than batch size");
        }
        auto mols = unflatten_mols(mol_atoms, mol_offsets, n_mols);
        std::vector<int> lig(ligand_idxs, ligand_idxs + n_ligand);
        std::vector<double> p(params, params + n_params);
        MoverPtr mv;
        if (precision == TMB_F32) {
            mv = std::make_shared<TIBDExchangeMove<float>>(N, lig, mols, p, temperature, nb_beta, cutoff, radius, seed, num_proposals_per_move, interval, batch_size);
        } else {
            mv = std::make_shared<TIBDExchangeMove<double>>(N, lig, mols, p, temperature, nb_beta, cutoff, radius, seed, num_proposals_per_move, interval, batch_size);
        }
        *out = new MoverPtr(mv);
    });
}
int tmb_inner_and_outer_mols(
    int precision, const int *center_atoms, int n_center, int N, const double *coords, const double *box, const int *mol_atoms,
    const int *mol_offsets, int n_mols, double radius, int *flags) {
    return guarded([&] {
        check_precision(precision);
        auto mols = unflatten_mols(mol_atoms, mol_offsets, n_mols);
        std::vector<int> c(center_atoms, center_atoms + (n_center > 0 ? n_center : 0));
        std::array<std::vector<int>, 2> r;
        if (precision == TMB_F32) {
            r = inner_and_outer_mols<float>(c, N, coords, box, mols, static_cast<float>(radius));
        } else {
            r = inner_and_outer_mols<double>(c, N, coords, box, mols, radius);
        }
        for (int m = 0; m < n_mols; m++) {
            flags[m] = 0;
        }
        for (int m : r[0]) {
            flags[m] = 1;
        }
    });
}
int tmb_translations_inside_and_outside_sphere(
    int precision, int n_translations, const double *box, const double *center, double radius, int seed, double *out) {
    return guarded([&] {
        check_precision(precision);
        if (precision == TMB_F32) {
            const float c[3] = {static_cast<float>(center[0]), static_cast<float>(center[1]), static_cast<float>(center[2])};
            widen(translations_inside_and_outside_sphere_host<float>(n_translations, box, c, static_cast<float>(radius), seed), out);
        } else {
            widen(translations_inside_and_outside_sphere_host<double>(n_translations, box, center, radius, seed), out);
        }
    });
}
int tmb_bd_exchange_move_num_target_mols(tmb_mover m, int *out) {
    return guarded([&] { *out = WITH_BD(m, mv.num_target_mols()); });
}
int tmb_bd_exchange_move_batch_size(tmb_mover m, int *out) {
    return guarded([&] { *out = static_cast<int>(WITH_BD(m, mv.batch_size())); });
}
int tmb_bd_exchange_move_initial_log_weights(tmb_mover m, int N, const double *coords, const double *box, double *out) {
    return guarded([&] {
        with_bd(
            m, [&](BDExchangeMove<float> &mv) { widen(mv.compute_initial_log_weights_host(N, coords, box), out); },
            [&](BDExchangeMove<double> &mv) { widen(mv.compute_initial_log_weights_host(N, coords, box), out); });
    });
}
extern "C++" {
template <typename Real>
static void incremental_weights(
    BDExchangeMove<Real> &mv, int N, const double *coords, const double *box, const int *mol_idxs, const double *quaternions,
    const double *translations, double *out) {
    const size_t B = mv.batch_size();
    std::vector<Real> q(B * 4), t(B * 3);
    for (size_t i = 0; i < q.size(); i++) {
        q[i] = static_cast<Real>(quaternions[i]);
    }
    for (size_t i = 0; i < t.size(); i++) {
        t[i] = static_cast<Real>(translations[i]);
    }
    auto w = mv.compute_incremental_log_weights_host(N, coords, box, mol_idxs, q.data(), t.data());
    for (size_t b = 0; b < w.size(); b++) {
        widen(w[b], out + b * w[b].size());
    }
}
} // extern "C++"
int tmb_bd_exchange_move_incremental_log_weights(
    tmb_mover m, int N, const double *coords, const double *box, const int *mol_idxs, const double *quaternions,
    const double *translations, double *out) {
    return guarded([&] {
        with_bd(
            m, [&](BDExchangeMove<float> &mv) { incremental_weights(mv, N, coords, box, mol_idxs, quaternions, translations, out); },
            [&](BDExchangeMove<double> &mv) { incremental_weights(mv, N, coords, box, mol_idxs, quaternions, translations, out); });
    });
}
int tmb_bd_exchange_move_get_params(tmb_mover m, double *out, int n_params) {
    return guarded([&] {
        std::vector<double> p = WITH_BD(m, mv.get_params());
        if (static_cast<int>(p.size()) != n_params) {
            throw std::runtime_error("number of params don't match");
        }
        std::memcpy(out, p.data(), sizeof(double) * p.size());
    });
}
int tmb_bd_exchange_move_set_params(tmb_mover m, const double *params, int n_params) {
    return guarded([&] {
        std::vector<double> p(params, params + (n_params > 0 ? n_params : 0));
        with_bd(
            m, [&](BDExchangeMove<float> &mv) { mv.set_params(p); }, [&](BDExchangeMove<double> &mv) { mv.set_params(p); });
    });
}
int tmb_bd_exchange_move_last_log_probability(tmb_mover m, double *out) {
    return guarded([&] { *out = WITH_BD(m, mv.log_probability_host()); });
}
int tmb_bd_exchange_move_last_raw_log_probability(tmb_mover m, double *out) {
    return guarded([&] { *out = WITH_BD(m, mv.raw_log_probability_host()); });
}
int tmb_bd_exchange_move_n_accepted(tmb_mover m, unsigned long long *out) {
    return guarded([&] { *out = WITH_BD(m, mv.n_accepted()); });
}
int tmb_bd_exchange_move_n_proposed(tmb_mover m, unsigned long long *out) {
    return guarded([&] { *out = WITH_BD(m, mv.n_proposed()); });
}
int tmb_bd_exchange_move_before_log_weights(tmb_mover m, double *out) {
    return guarded([&] {
        with_bd(
            m, [&](BDExchangeMove<float> &mv) { widen(mv.get_before_log_weights(), out); },
            [&](BDExchangeMove<double> &mv) { widen(mv.get_before_log_weights(), out); });
    });
}
int tmb_bd_exchange_move_after_log_weights(tmb_mover m, double *out) {
    return guarded([&] {
        with_bd(
            m, [&](BDExchangeMove<float> &mv) { widen(mv.get_after_log_weights(), out); },
            [&](BDExchangeMove<double> &mv) { widen(mv.get_after_log_weights(), out); });
    });
}

extern "C++" {
template <typename Real>
static void mol_energies(
    int N, const std::vector<std::vector<int>> &mols, double beta, double cutoff, const double *coords, const double *params,
    const double *box, tmb_i128 *out) {
    NonbondedMolEnergyPotential<Real> pot(N, mols, beta, cutoff);
    if (coords == nullptr) {
        return;
    }
    std::vector<i128> e = pot.mol_energies_host(N, N * P_PER_ATOM, coords, params, box);
    std::memcpy(out, e.data(), sizeof(i128) * e.size());
}
} // extern "C++"
int tmb_nonbonded_mol_energies(
    int precision, int N, const int *mol_atoms, const int *mol_offsets, int n_mols, double beta, double cutoff,
    const double *coords, const double *params, const double *box, tmb_i128 *out) {
    return guarded([&] {
        check_precision(precision);
        auto mols = unflatten_mols(mol_atoms, mol_offsets, n_mols);
        if (precision == TMB_F32) {
            mol_energies<float>(N, mols, beta, cutoff, coords, params, box, out);
        } else {
            mol_energies<double>(N, mols, beta, cutoff, coords, params, box, out);
        }
    });
}
int tmb_atom_by_atom_energies(
    int precision, int N, const int *target_atoms, int T, const double *coords, const double *params, const double *box,
    double nb_beta, double cutoff, double *out) {
    return guarded([&] {
        check_precision(precision);
        std::vector<int> t(target_atoms, target_atoms + (T > 0 ? T : 0));
        if (precision == TMB_F32) {
            widen(atom_by_atom_energies<float>(N, t, coords, params, box, static_cast<float>(nb_beta), static_cast<float>(cutoff)), out);
        } else {
            widen(atom_by_atom_energies<double>(N, t, coords, params, box, nb_beta, cutoff), out);
        }
    });
}

extern "C++" {
template <typename Real> static std::vector<std::vector<Real>> unflatten_values(const double *values, const int *offsets, int num_segments) {
    std::vector<std::vector<Real>> v(num_segments > 0 ? num_segments : 0);
    for (int s = 0; s < num_segments; s++) {
        for (int k = offsets[s]; k < offsets[s + 1]; k++) {
            v[s].push_back(static_cast<Real>(values[k]));
        }
    }
    return v;
}
} // extern "C++"
int tmb_segmented_logsumexp(
    int precision, int max_vals_per_segment, int num_segments_max, const double *values, const int *offsets, int num_segments,
    double *out) {
    return guarded([&] {
        check_precision(precision);
        if (precision == TMB_F32) {
            SegmentedSumExp<float> s(max_vals_per_segment, num_segments_max);
            auto v = unflatten_values<float>(values, offsets, num_segments);
            widen(s.logsumexp_host(v), out);
        } else {
            SegmentedSumExp<double> s(max_vals_per_segment, num_segments_max);
            auto v = unflatten_values<double>(values, offsets, num_segments);
            widen(s.logsumexp_host(v), out);
        }
    });
}

extern "C++" {
struct SamplerHandle {
    std::unique_ptr<SegmentedWeightedRandomSampler<float>> f32;
    std::unique_ptr<SegmentedWeightedRandomSampler<double>> f64;
};
} // extern "C++"
int tmb_weighted_sampler_create(int precision, int max_vals_per_segment, int num_segments, int seed, tmb_sampler *out) {
    return guarded([&] {
        check_precision(precision);
        auto h = std::make_unique<SamplerHandle>();
        if (precision == TMB_F32) {
            h->f32 = std::make_unique<SegmentedWeightedRandomSampler<float>>(max_vals_per_segment, num_segments, seed);
        } else {
            h->f64 = std::make_unique<SegmentedWeightedRandomSampler<double>>(max_vals_per_segment, num_segments, seed);
        }
        *out = h.release();
    });
}
int tmb_weighted_sampler_destroy(tmb_sampler s) {
    return guarded([&] { delete static_cast<SamplerHandle *>(s); });
}
int tmb_weighted_sampler_sample(tmb_sampler s, const double *weights, const int *offsets, int num_segments, int *out) {
    return guarded([&] {
        if (s == nullptr) {
            throw std::runtime_error("null sampler handle");
        }
        SamplerHandle &h = *static_cast<SamplerHandle *>(s);
        std::vector<int> r;
        if (h.f32) {
            r = h.f32->sample_host(unflatten_values<float>(weights, offsets, num_segments));
        } else {
            r = h.f64->sample_host(unflatten_values<double>(weights, offsets, num_segments));
        }
        std::memcpy(out, r.data(), sizeof(int) * r.size());
    });
}
int tmb_rotate_coords(int precision, int N, int n_rotations, const double *coords, const double *quaternions, double *out) {
    return guarded([&] {
        check_precision(precision);
        if (precision == TMB_F32) {
            std::vector<float> q(quaternions, quaternions + static_cast<size_t>(n_rotations) * 4);
            rotate_coordinates_host<float>(N, n_rotations, coords, q.data(), out);
        } else {
            rotate_coordinates_host<double>(N, n_rotations, coords, quaternions, out);
        }
    });
}
int tmb_rotate_and_translate_mol(
    int precision, int N, int batch_size, const double *coords, const double *box, const double *quaternions,
    const double *translations, double *out) {
    return guarded([&] {
        check_precision(precision);
        if (precision == TMB_F32) {
            std::vector<float> q(quaternions, quaternions + static_cast<size_t>(batch_size) * 4);
            std::vector<float> t(translations, translations + static_cast<size_t>(batch_size) * 3);
            rotate_and_translate_mol_host<float>(N, batch_size, coords, box, q.data(), t.data(), out);
        } else {
            rotate_and_translate_mol_host<double>(N, batch_size, coords, box, quaternions, translations, out);
        }
    });
}

int tmb_mover_destroy(tmb_mover m) {
    return guarded([&] { delete static_cast<MoverPtr *>(m); });
}
int tmb_mover_set_interval(tmb_mover m, int interval) {
    return guarded([&] { as_mover(m)->set_interval(interval); });
}
int tmb_mover_get_interval(tmb_mover m, int *out) {
    return guarded([&] { *out = as_mover(m)->get_interval(); });
}
int tmb_mover_set_step(tmb_mover m, int step) {
    return guarded([&] { as_mover(m)->set_step(step); });
}
int tmb_mover_move_host(tmb_mover m, int N, const double *coords, const double *box, double *out_coords, double *out_box) {
    return guarded([&] {
        auto r = as_mover(m)->move_host(N, coords, box);
        std::memcpy(out_coords, r[0].data(), r[0].size() * sizeof(double));
        std::memcpy(out_box, r[1].data(), r[1].size() * sizeof(double));
    });
}
int tmb_barostat_set_volume_scale_factor(tmb_mover m, double v) {
    return guarded([&] { as_barostat(m)->set_volume_scale_factor(v); });
}
int tmb_barostat_get_volume_scale_factor(tmb_mover m, double *out) {
    return guarded([&] { *out = as_barostat(m)->get_volume_scale_factor(); });
}
int tmb_barostat_set_adaptive_scaling(tmb_mover m, int on) {
    return guarded([&] { as_barostat(m)->set_adaptive_scaling(on != 0); });
}
int tmb_barostat_get_adaptive_scaling(tmb_mover m, int *out) {
    return guarded([&] { *out = as_barostat(m)->get_adaptive_scaling() ? 1 : 0; });
}
int tmb_barostat_set_pressure(tmb_mover m, double pressure) {
    return guarded([&] { as_barostat(m)->set_pressure(pressure); });
}
int tmb_barostat_last_uniforms(tmb_mover m, float *out2) {
    return guarded([&] {
        auto u = as_barostat(m)->last_uniforms();
        out2[0] = u[0];
        out2[1] = u[1];
    });
}
int tmb_barostat_counters(tmb_mover m, int *out2) {
    return guarded([&] {
        auto c = as_barostat(m)->counters();
        out2[0] = c[0];
        out2[1] = c[1];
    });
}
int tmb_context_destroy(tmb_context ctx) {
    return guarded([&] { delete static_cast<Context *>(ctx); });
}
int tmb_context_step(tmb_context ctx) {
    return guarded([&] { as_ctx(ctx).step(); });
}
int tmb_context_initialize(tmb_context ctx) {
    return guarded([&] { as_ctx(ctx).initialize(); });
}
int tmb_context_finalize(tmb_context ctx) {
    return guarded([&] { as_ctx(ctx).finalize(); });
}
int tmb_context_multiple_steps(tmb_context ctx, int n_steps, int n_samples, double *h_x, double *h_box) {
    return guarded([&] { as_ctx(ctx).multiple_steps(n_steps, n_samples, h_x, h_box); });
}
int tmb_context_setup_local_md(tmb_context ctx, double temperature, int freeze_reference) {
    return guarded([&] { as_ctx(ctx).setup_local_md(temperature, freeze_reference != 0); });
}
int tmb_context_multiple_steps_local(
    tmb_context ctx, int n_steps, const int *local_idxs, int n_local_idxs, int n_samples, double radius, double k, int seed,
    double *h_x, double *h_box) {
    return guarded([&] {
        as_ctx(ctx).multiple_steps_local(
            n_steps, std::vector<int>(local_idxs, local_idxs + n_local_idxs), n_samples, radius, k, seed, h_x, h_box);
    });
}
int tmb_context_multiple_steps_local_selection(
    tmb_context ctx, int n_steps, int reference_idx, const int *selection_idxs, int n_selection_idxs, int n_samples,
    double radius, double k, double *h_x, double *h_box) {
    return guarded([&] {
        as_ctx(ctx).multiple_steps_local_selection(
            n_steps, reference_idx, std::vector<int>(selection_idxs, selection_idxs + n_selection_idxs), n_samples, radius, k,
            h_x, h_box);
    });
}
int tmb_context_local_md_free_idxs(tmb_context ctx, unsigned int *out) {
    return guarded([&] {
        const LocalMD *l = as_ctx(ctx).local_md();
        if (l == nullptr) {
            throw std::runtime_error("local md has not been set up");
        }
        std::copy(l->selected_host().begin(), l->selected_host().end(), out);
    });
}
int tmb_context_set_x_t(tmb_context ctx, const double *x) {
    return guarded([&] { as_ctx(ctx).set_x_t(x); });
}
int tmb_context_set_v_t(tmb_context ctx, const double *v) {
    return guarded([&] { as_ctx(ctx).set_v_t(v); });
}
int tmb_context_set_box(tmb_context ctx, const double *box) {
    return guarded([&] { as_ctx(ctx).set_box(box); });
}
int tmb_context_get_x_t(tmb_context ctx, double *x) {
    return guarded([&] { as_ctx(ctx).get_x_t(x); });
}
int tmb_context_get_v_t(tmb_context ctx, double *v) {
    return guarded([&] { as_ctx(ctx).get_v_t(v); });
}
int tmb_context_get_box(tmb_context ctx, double *box) {
    return guarded([&] { as_ctx(ctx).get_box(box); });
}
int tmb_context_num_atoms(tmb_context ctx, int *out) {
    return guarded([&] { *out = as_ctx(ctx).num_atoms(); });
}
int tmb_context_set_stream(tmb_context ctx, void *cuda_stream) {
    return guarded([&] { as_ctx(ctx).set_stream(static_cast<cudaStream_t>(cuda_stream)); });
}
int tmb_context_set_use_graphs(tmb_context ctx, int on) {
    return guarded([&] { as_ctx(ctx).set_use_graphs(on != 0); });
}
int tmb_context_device_state(tmb_context ctx, double **d_x, double **d_v, double **d_box) {
    return guarded([&] {
        *d_x = as_ctx(ctx).d_x();
        *d_v = as_ctx(ctx).d_v();
        *d_box = as_ctx(ctx).d_box();
    });
}

// ---- neighbour list / hilbert ---------------------------------------------------------------------------------------
int tmb_neighborlist_create(int precision, int N, tmb_neighborlist *out) {
    return guarded([&] {
        check_precision(precision);
        std::unique_ptr<NeighborlistHandle> h(new NeighborlistHandle());
        h->precision = precision;
        if (precision == TMB_F32) {
            h->f32.reset(new Neighborlist<float>(N));
        } else {
            h->f64.reset(new Neighborlist<double>(N));
        }
        *out = h.release();
    });
}
int tmb_neighborlist_destroy(tmb_neighborlist nb) {
    return guarded([&] { delete static_cast<NeighborlistHandle *>(nb); });
}

#define NB_DISPATCH(h, expr_f32, expr_f64)                                                                             \
    do {                                                                                                               \
        if ((h).precision == TMB_F32) {                                                                                \
            expr_f32;                                                                                                  \
        } else {                                                                                                       \
            expr_f64;                                                                                                  \
        }                                                                                                              \
    } while (0)

int tmb_neighborlist_build(
    tmb_neighborlist nb, int N, const double *coords, const double *box, double cutoff, int *n_row_blocks, int *n_entries) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, h.last = h.f32->get_nblist_host(N, coords, box, cutoff), h.last = h.f64->get_nblist_host(N, coords, box, cutoff));
        size_t total = 0;
        for (auto &row : h.last) {
            total += row.size();
        }
        *n_row_blocks = static_cast<int>(h.last.size());
        *n_entries = static_cast<int>(total);
    });
}
int tmb_neighborlist_fetch(tmb_neighborlist nb, int32_t *offsets, int32_t *atoms) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        int32_t off = 0;
        for (size_t r = 0; r < h.last.size(); r++) {
            offsets[r] = off;
            for (int a : h.last[r]) {
                atoms[off++] = a;
            }
        }
        offsets[h.last.size()] = off;
    });
}
int tmb_neighborlist_compute_block_bounds(
    tmb_neighborlist nb, int N, const double *coords, const double *box, double *ctrs, double *exts) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, h.f32->compute_block_bounds_host(N, coords, box, ctrs, exts), h.f64->compute_block_bounds_host(N, coords, box, ctrs, exts));
    });
}
int tmb_neighborlist_set_row_idxs(tmb_neighborlist nb, const uint32_t *idxs, int n) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        std::vector<unsigned int> v(idxs, idxs + (n > 0 ? n : 0));
        NB_DISPATCH(h, h.f32->set_row_idxs(v), h.f64->set_row_idxs(v));
    });
}
int tmb_neighborlist_reset_row_idxs(tmb_neighborlist nb) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, h.f32->reset_row_idxs(), h.f64->reset_row_idxs());
    });
}
int tmb_neighborlist_resize(tmb_neighborlist nb, int size) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, h.f32->resize(size), h.f64->resize(size));
    });
}
int tmb_neighborlist_get_tile_ixn_count(tmb_neighborlist nb, unsigned int *out) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, *out = h.f32->num_tile_ixns(), *out = h.f64->num_tile_ixns());
    });
}
int tmb_neighborlist_get_max_ixn_count(tmb_neighborlist nb, int *out) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, *out = h.f32->max_ixn_count(), *out = h.f64->max_ixn_count());
    });
}
int tmb_neighborlist_get_num_row_idxs(tmb_neighborlist nb, int *out) {
    return guarded([&] {
        NeighborlistHandle &h = as_nb(nb);
        NB_DISPATCH(h, *out = h.f32->get_num_row_idxs(), *out = h.f64->get_num_row_idxs());
    });
}

int tmb_hilbert_sort_create(int size, tmb_hilbert_sort *out) {
    return guarded([&] { *out = new HilbertSort(size); });
}
int tmb_hilbert_sort_destroy(tmb_hilbert_sort hs) {
    return guarded([&] { delete static_cast<HilbertSort *>(hs); });
}
int tmb_hilbert_sort_sort(tmb_hilbert_sort hs, int N, const double *coords, const double *box, uint32_t *perm) {
    return guarded([&] {
        if (hs == nullptr) {
            throw std::runtime_error("null hilbert sort handle");
        }
        std::vector<unsigned int> p = static_cast<HilbertSort *>(hs)->sort_host(N, coords, box);
        std::copy(p.begin(), p.end(), perm);
    });
}

int tmb_fill_normal(float *out, int n_atoms, uint64_t seed, uint64_t step) {
    return guarded([&] {
        DeviceBuffer<float> d(static_cast<size_t>(n_atoms) * 3);
        launch_fill_normal(d.data, n_atoms, seed, step, main_stream());
        TMB_CUDA(cudaStreamSynchronize(main_stream()));
        d.copy_to(out);
    });
}

// ---- HREX host loop (reference timemachine/md/hrex.py:50-129, `_run_neighbor_swaps`) ---------------------------------
int tmb_hrex_run_neighbor_swaps(
    int n_states, int n_replicas, int n_pairs, const int32_t *neighbor_pairs, const double *log_q_kl, int n_attempts,
    const int32_t *pair_idxs, const double *uniform_samples, int32_t *replica_idx_by_state, uint32_t *proposed,
    uint32_t *accepted) {
    return guarded([&] {
        for (int p = 0; p < n_pairs; p++) {
            proposed[p] = 0;
            accepted[p] = 0;
            for (int k = 0; k < 2; k++) {
                const int s = neighbor_pairs[p * 2 + k];
                if (s < 0 || s >= n_states) {
                    throw std::runtime_error("neighbor_pairs holds a state outside [0, n_states)");
                }
            }
        }
        for (int s = 0; s < n_states; s++) {
            if (replica_idx_by_state[s] < 0 || replica_idx_by_state[s] >= n_replicas) {
                throw std::runtime_error("replica_idx_by_state holds a replica outside [0, n_replicas)");
            }
        }
        for (int i = 0; i < n_attempts; i++) {
            const int p = pair_idxs[i];
            if (p < 0 || p >= n_pairs) {
                throw std::runtime_error("pair_idxs holds an index outside [0, n_pairs)");
            }
            const int s_a = neighbor_pairs[p * 2 + 0], s_b = neighbor_pairs[p * 2 + 1];
            proposed[p] += 1;
            const int r_a = replica_idx_by_state[s_a], r_b = replica_idx_by_state[s_b];
            const double *qa = log_q_kl + static_cast<size_t>(r_a) * n_states;
            const double *qb = log_q_kl + static_cast<size_t>(r_b) * n_states;
            const double diff = (qa[s_b] + qb[s_a]) - (qa[s_a] + qb[s_b]);
            const double m = (0.0 < diff) ? 0.0 : diff; // min(diff, 0) that lets NaN through, like jnp.minimum
            if (uniform_samples[i] < std::exp(m)) {
                replica_idx_by_state[s_a] = r_b;
                replica_idx_by_state[s_b] = r_a;
                accepted[p] += 1;
            }
        }
    });
}

} // extern "C"
