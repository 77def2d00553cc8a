// A HOOMD-blue component (pybind11 module, built against the reference's own headers like
// example_plugins/pair_plugin/module.cc) that puts libhoomd_b200.so behind the reference's ForceCompute
// interface: hoomd.md.Integrator(forces=[hoomd_b200_plugin.pair.LJ(...)]) runs the UNMODIFIED reference —
// its integrator, its methods, its writers — with the pair force evaluated by this repository's CUDA path.
//
// Boundary. The reference build in this image is CPU-only (no GPU arrays to hand over), so the component
// drives layer 2 of include/hoomd_b200.h: positions go in through hb2_system_set_particles (tag order),
// the force, energy and virial arrays come back through hb2_pair_get_* and are scattered to the reference's
// index order with its rtag table. In a GPU build of the reference the same class would pass the device
// pointers of ParticleData to the layer-1 drivers. The library is bound at run time
// (dlopen): the module has no link-time dependency on CUDA.
//
// Replaces hoomd::md::PotentialPair<EvaluatorPairLJ>::computeForces (hoomd/md/PotentialPair.h:587-794)
// together with the neighbour list it would have asked (hoomd/md/NeighborListBinned.cc:36-179).
#include "hoomd/ForceCompute.h"
#include "hoomd/ParticleData.h"
#include "hoomd/SystemDefinition.h"

#include <dlfcn.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace hoomd
    {
namespace b200
    {
// errors are raised as Python RuntimeError directly (the message of the library's hb2_last_error included)
[[noreturn]] static void fail(const std::string& msg)
    {
    PyErr_SetString(PyExc_RuntimeError, msg.c_str());
    throw pybind11::error_already_set();
    }
// the part of include/hoomd_b200.h this component binds (plain C ABI)
struct Api
    {
    void* handle = nullptr;
    int (*system_create)(void**, int, uint32_t, uint32_t) = nullptr;
    void (*system_destroy)(void*) = nullptr;
    int (*system_set_box)(void*, double, double, double, double, double, double) = nullptr;
    int (*system_set_particles)(void*, uint32_t, const float*, const float*, const float*, const uint32_t*,
                                const int32_t*)
        = nullptr;
    int (*system_set_compute_virial)(void*, int) = nullptr;
    int (*nlist_create)(void**, void*, double, uint32_t, int) = nullptr;
    void (*nlist_destroy)(void*) = nullptr;
    uint64_t (*nlist_num_builds)(const void*) = nullptr;
    int (*pair_create)(void**, void*, void*, int, int) = nullptr;
    void (*pair_destroy)(void*) = nullptr;
    int (*pair_set_params)(void*, uint32_t, uint32_t, const double*, uint32_t) = nullptr;
    int (*pair_set_r_cut)(void*, uint32_t, uint32_t, double) = nullptr;
    int (*pair_compute)(void*, uint64_t) = nullptr;
    int (*pair_get_forces)(void*, double*) = nullptr;
    int (*pair_get_energies)(void*, double*) = nullptr;
    int (*pair_get_virials)(void*, double*) = nullptr;
    const char* (*last_error)() = nullptr;
    uint64_t (*launch_count)() = nullptr;

    void load(const std::string& path)
        {
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle)
            fail(std::string("hoomd_b200_plugin: cannot load ") + path + ": " + dlerror());
#define HB2_SYM(field, name)                                                              \
    *(void**)(&field) = dlsym(handle, name);                                              \
    if (!field)                                                                           \
        fail(std::string("hoomd_b200_plugin: missing symbol ") + name);
        HB2_SYM(system_create, "hb2_system_create")
        HB2_SYM(system_destroy, "hb2_system_destroy")
        HB2_SYM(system_set_box, "hb2_system_set_box")
        HB2_SYM(system_set_particles, "hb2_system_set_particles")
        HB2_SYM(system_set_compute_virial, "hb2_system_set_compute_virial")
        HB2_SYM(nlist_create, "hb2_nlist_create")
        HB2_SYM(nlist_destroy, "hb2_nlist_destroy")
        HB2_SYM(nlist_num_builds, "hb2_nlist_num_builds")
        HB2_SYM(pair_create, "hb2_pair_create")
        HB2_SYM(pair_destroy, "hb2_pair_destroy")
        HB2_SYM(pair_set_params, "hb2_pair_set_params")
        HB2_SYM(pair_set_r_cut, "hb2_pair_set_r_cut")
        HB2_SYM(pair_compute, "hb2_pair_compute")
        HB2_SYM(pair_get_forces, "hb2_pair_get_forces")
        HB2_SYM(pair_get_energies, "hb2_pair_get_energies")
        HB2_SYM(pair_get_virials, "hb2_pair_get_virials")
        HB2_SYM(last_error, "hb2_last_error")
        HB2_SYM(launch_count, "hb2_launch_count")
#undef HB2_SYM
        }
    };

class PairLJB200 : public ForceCompute
    {
    public:
    PairLJB200(std::shared_ptr<SystemDefinition> sysdef, const std::string& library, int device, double buffer,
               int shift_mode)
        : ForceCompute(sysdef), m_buffer(buffer), m_mode(shift_mode)
        {
        m_api.load(library);
        const unsigned int nt = m_pdata->getNTypes();
        check(m_api.system_create(&m_sys, device, nt, sysdef->getSeed()), "hb2_system_create");
        check(m_api.nlist_create(&m_nl, m_sys, buffer, 1, 1), "hb2_nlist_create");
        check(m_api.pair_create(&m_pair, m_sys, m_nl, 0 /* HB2_EVAL_LJ */, shift_mode), "hb2_pair_create");
        }

    virtual ~PairLJB200()
        {
        if (m_pair)
            m_api.pair_destroy(m_pair);
        if (m_nl)
            m_api.nlist_destroy(m_nl);
        if (m_sys)
            m_api.system_destroy(m_sys);
        }

    // hoomd.md.pair.LJ.params[(a, b)] = dict(epsilon, sigma); r_cut[(a, b)]
    void setParams(unsigned int ti, unsigned int tj, double epsilon, double sigma, double r_cut)
        {
        const double user[2] = {epsilon, sigma};
        check(m_api.pair_set_params(m_pair, ti, tj, user, 2), "hb2_pair_set_params");
        check(m_api.pair_set_r_cut(m_pair, ti, tj, r_cut), "hb2_pair_set_r_cut");
        }

    uint64_t getNumBuilds()
        {
        return m_api.nlist_num_builds(m_nl);
        }
    uint64_t getLaunchCount()
        {
        return m_api.launch_count();
        }

    protected:
    virtual void computeForces(uint64_t timestep)
        {
        const unsigned int N = m_pdata->getN();
        if (m_sysdef->isDomainDecomposed())
            fail("hoomd_b200_plugin: one rank only");
        const BoxDim box = m_pdata->getGlobalBox();
        const Scalar3 L = box.getL();
        if (box.getTiltFactorXY() != 0 || box.getTiltFactorXZ() != 0 || box.getTiltFactorYZ() != 0)
            fail("hoomd_b200_plugin: orthorhombic boxes only");
        const PDataFlags flags = m_pdata->getFlags();
        const bool want_virial = flags[pdata_flag::pressure_tensor];
        // ---- state in, tag order
        m_pos.resize(3 * (size_t)N);
        m_vel.assign(3 * (size_t)N, 0.f);
        m_mass.assign(N, 1.f);
        m_type.resize(N);
            {
            ArrayHandle<Scalar4> h_pos(m_pdata->getPositions(), access_location::host, access_mode::read);
            ArrayHandle<unsigned int> h_tag(m_pdata->getTags(), access_location::host, access_mode::read);
            for (unsigned int i = 0; i < N; i++)
                {
                const unsigned int t = h_tag.data[i];
                m_pos[3 * (size_t)t] = (float)h_pos.data[i].x;
                m_pos[3 * (size_t)t + 1] = (float)h_pos.data[i].y;
                m_pos[3 * (size_t)t + 2] = (float)h_pos.data[i].z;
                m_type[t] = (uint32_t)__scalar_as_int(h_pos.data[i].w);
                }
            }
        if (L.x != m_L[0] || L.y != m_L[1] || L.z != m_L[2])
            {
            check(m_api.system_set_box(m_sys, L.x, L.y, L.z, 0, 0, 0), "hb2_system_set_box");
            m_L[0] = L.x;
            m_L[1] = L.y;
            m_L[2] = L.z;
            }
        check(m_api.system_set_compute_virial(m_sys, want_virial ? 1 : 0), "hb2_system_set_compute_virial");
        check(m_api.system_set_particles(m_sys, N, m_pos.data(), m_vel.data(), m_mass.data(), m_type.data(), nullptr),
              "hb2_system_set_particles");
        // ---- cell list, neighbour list, pair forces on the GPU
        check(m_api.pair_compute(m_pair, timestep), "hb2_pair_compute");
        // ---- results out, tag order -> the reference's index order
        m_f.resize(3 * (size_t)N);
        m_e.resize(N);
        check(m_api.pair_get_forces(m_pair, m_f.data()), "hb2_pair_get_forces");
        check(m_api.pair_get_energies(m_pair, m_e.data()), "hb2_pair_get_energies");
        if (want_virial)
            {
            m_w.resize(6 * (size_t)N);
            check(m_api.pair_get_virials(m_pair, m_w.data()), "hb2_pair_get_virials");
            }
        ArrayHandle<Scalar4> h_force(m_force, access_location::host, access_mode::overwrite);
        ArrayHandle<Scalar> h_virial(m_virial, access_location::host, access_mode::overwrite);
        ArrayHandle<unsigned int> h_tag(m_pdata->getTags(), access_location::host, access_mode::read);
        for (unsigned int i = 0; i < N; i++)
            {
            const size_t t = h_tag.data[i];
            h_force.data[i] = make_scalar4((Scalar)m_f[3 * t], (Scalar)m_f[3 * t + 1], (Scalar)m_f[3 * t + 2],
                                           (Scalar)m_e[t]);
            for (int q = 0; q < 6; q++)
                h_virial.data[q * m_virial_pitch + i] = want_virial ? (Scalar)m_w[6 * t + q] : Scalar(0);
            }
        }

    private:
    void check(int status, const char* what)
        {
        if (status != 0)
            fail(std::string("hoomd_b200_plugin: ") + what + ": " + (m_api.last_error ? m_api.last_error() : "?"));
        }
    Api m_api;
    void* m_sys = nullptr;
    void* m_nl = nullptr;
    void* m_pair = nullptr;
    double m_buffer;
    int m_mode;
    double m_L[3] = {0, 0, 0};
    std::vector<float> m_pos, m_vel, m_mass;
    std::vector<uint32_t> m_type;
    std::vector<double> m_f, m_e, m_w;
    };

PYBIND11_MODULE(_b200_plugin, m)
    {
    pybind11::class_<PairLJB200, ForceCompute, std::shared_ptr<PairLJB200>>(m, "PairLJB200")
        .def(pybind11::init<std::shared_ptr<SystemDefinition>, const std::string&, int, double, int>())
        .def("setParams", &PairLJB200::setParams)
        .def("getNumBuilds", &PairLJB200::getNumBuilds)
        .def("getLaunchCount", &PairLJB200::getLaunchCount);
    }
    } // namespace b200
    } // namespace hoomd
