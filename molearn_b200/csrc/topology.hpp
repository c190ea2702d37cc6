// Host-side topology container shared by topology.cpp (builder) and energy.cu (device upload).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct mlp_topology {
    int32_t n_atoms = 0;
    int32_t P = 0;  // torsion series slots (max number of cosine terms of any torsion)
    std::vector<std::string> names, resnames, types;  // after normalisation
    std::vector<int32_t> resids;
    std::vector<int32_t> bonds;      // [nb,2]  (i2,i1) with i2<i1, reference order
    std::vector<int32_t> angles;     // [na,3]
    std::vector<int32_t> torsions;   // [nt,4]
    std::vector<int32_t> pairs14;    // [nt,2]
    std::vector<double> bond_para;   // [nb,2] (r0,K)
    std::vector<double> angle_para;  // [na,2] (theta0,K)
    std::vector<double> torsion_para;  // [nt,4,P] (factor,barrier,phase,|period|)
    std::vector<double> charges;     // [N]
    std::vector<float> vdw_R, vdw_e; // [N] fp32 like the reference's torch.tensor(...)
    std::vector<int32_t> excl;       // [ne,2] unique sorted pairs i<j
};

// thread-local error message helpers (capi)
void mlp_set_error(const std::string& msg);
