// KartLQR.hpp — C++ host-side mirror of the reference's C# namespace KartGame.AI.LQR over the C-ABI (include/hk_abi.h).
// The reference host language is C#, whose toolchain is absent from this image, so the compiled-host flavour of the
// drop-in is C++ (the C# shim a maintainer would use is csharp/KartLQR.cs).  Same names, argument meaning and error
// behaviour as
//   KartLQR.solveFeedbackLQR            Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17
//   KartLQRDynamics / LinearizedBicycle Assets/Karting/Scripts/AI/LQR/KartLQRDynamics.cs:14-73
//   KartLQRCosts / LQRCheckpointReachAvoidCost   Assets/Karting/Scripts/AI/LQR/KartLQRCosts.cs:13-141
// Providers only describe a problem; every solve runs in libhk_b200 (CUDA). Dimension mismatches throw
// std::invalid_argument where MathNet would throw ArgumentException.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../include/hk_abi.h"

namespace KartGame { namespace AI { namespace LQR {

using Vector = std::vector<double>;
struct Matrix {                                   // dense row-major stand-in for MathNet Matrix<double>
    int rows = 0, cols = 0;
    std::vector<double> a;
    Matrix() = default;
    Matrix(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
    double& operator()(int r, int c) { return a[(size_t)r * cols + c]; }
    double operator()(int r, int c) const { return a[(size_t)r * cols + c]; }
    static Matrix Identity(int n) { Matrix m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
};

namespace KartMPC { enum { xIndex = 0, zIndex = 1, vIndex = 2, hIndex = 3 }; }   // MPC/KartMPC.cs:15-18

inline void hk_check(int status)
{
    if (status == HK_OK) return;
    std::string msg = hk_last_error();
    if (status == HK_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error("hk_b200 status " + std::to_string(status) + ": " + msg);
}

class KartLQRDynamics {                           // KartLQRDynamics.cs:14-20
public:
    virtual ~KartLQRDynamics() = default;
    virtual const Matrix& getA() = 0;
    virtual const Matrix& getB() = 0;
    virtual int getXDim() const = 0;
    virtual int getUDim() const = 0;
};

class LinearizedBicycle : public KartLQRDynamics { // KartLQRDynamics.cs:25-73
public:
    static constexpr int xDim = 4, uDim = 2;
    LinearizedBicycle(double dt, const Vector& initial) : dt_(dt), initial_(initial) {}
    const Matrix& getA() override
    {
        if (A_.rows == 0) {
            A_ = Matrix::Identity(xDim);
            A_(KartMPC::xIndex, KartMPC::vIndex) = std::cos(initial_[KartMPC::hIndex]) * dt_;
            A_(KartMPC::zIndex, KartMPC::vIndex) = std::sin(initial_[KartMPC::hIndex]) * dt_;
            A_(KartMPC::xIndex, KartMPC::hIndex) = -std::sin(initial_[KartMPC::hIndex]) * dt_ * initial_[KartMPC::vIndex];
            A_(KartMPC::zIndex, KartMPC::hIndex) = std::cos(initial_[KartMPC::hIndex]) * dt_ * initial_[KartMPC::vIndex];
        }
        return A_;
    }
    const Matrix& getB() override
    {
        if (B_.rows == 0) { B_ = Matrix(xDim, uDim); B_(KartMPC::vIndex, 0) = dt_; B_(KartMPC::hIndex, 1) = dt_; }
        return B_;
    }
    int getXDim() const override { return xDim; }
    int getUDim() const override { return uDim; }
    const Vector& initial() const { return initial_; }
    double dt() const { return dt_; }
private:
    double dt_; Vector initial_; Matrix A_, B_;
};

class KartLQRCosts {                              // KartLQRCosts.cs:13-19
public:
    virtual ~KartLQRCosts() = default;
    virtual const Vector& getQVec() = 0;
    virtual const Matrix& getQMatrix() = 0;
    virtual const Matrix& getRMatrix() = 0;
};

class LQRCheckpointReachAvoidCost : public KartLQRCosts {   // KartLQRCosts.cs:25-141
public:
    using Weights = std::map<int, double>;        // Dictionary<int,double>; x, z, v, h keys iterate in index order as in HKA:930-962
    LQRCheckpointReachAvoidCost(Vector targetState, Weights targetWeights, double controlWeight, KartLQRDynamics* currentDynamics,
                                std::vector<Vector> opponentTargetStates, std::vector<Weights> opponentTargetWeights,
                                std::map<int, std::vector<double>> avoidWeights, std::map<int, std::vector<int>> avoidIndices,
                                std::vector<KartLQRDynamics*> avoidDynamics)
        : target_(std::move(targetState)), tw_(std::move(targetWeights)), cw_(controlWeight), cur_(currentDynamics),
          otgt_(std::move(opponentTargetStates)), otw_(std::move(opponentTargetWeights)), aw_(std::move(avoidWeights)),
          ai_(std::move(avoidIndices)), avoid_(std::move(avoidDynamics)) {}

    const Matrix& getQMatrix() override                      // :57-98
    {
        if (Q_.rows == 0) {
            const int n = total();
            Q_ = Matrix(n, n);
            for (auto& kv : aw_) {
                const int s = kv.first;
                int curr = cur_->getXDim();
                double tot = 0.0;
                for (size_t i = 0; i < avoid_.size(); ++i) {
                    const int t = curr + ai_.at(s)[i];
                    const double w = kv.second[i];
                    Q_(s, t) = w; Q_(t, s) = w; Q_(t, t) = -w;
                    tot -= w;
                    curr += avoid_[i]->getXDim();
                }
                Q_(s, s) = tot;
            }
            for (auto& kv : tw_) Q_(kv.first, kv.first) += kv.second;
            int curr = cur_->getXDim();
            for (size_t i = 0; i < otw_.size(); ++i) {
                for (auto& kv : otw_[i]) Q_(curr + kv.first, curr + kv.first) = -kv.second;     // assignment (quirk Q4)
                curr += avoid_[i]->getXDim();
            }
        }
        return Q_;
    }
    const Vector& getQVec() override                         // :103-127
    {
        if (q_.empty()) {
            q_.assign(total(), 0.0);
            const int xd = cur_->getXDim();
            for (int s = 0; s < xd; ++s) q_[s] = -target_[s];
            for (auto& kv : tw_) q_[kv.first] = q_[kv.first] * kv.second;
            int curr = xd;
            for (size_t i = 0; i < otgt_.size(); ++i) {
                const int d = avoid_[i]->getXDim();
                for (int s = 0; s < d; ++s) q_[curr + s] = otgt_[i][s];
                for (auto& kv : otw_[i]) q_[curr + kv.first] = q_[curr + kv.first] * -kv.second;
                curr += d;
            }
        }
        return q_;
    }
    const Matrix& getRMatrix() override                      // :132-140
    {
        if (R_.rows == 0) { R_ = Matrix::Identity(cur_->getUDim()); for (auto& v : R_.a) v *= cw_; }
        return R_;
    }
private:
    int total() const { int n = cur_->getXDim(); for (auto* d : avoid_) n += d->getXDim(); return n; }
    Vector target_; Weights tw_; double cw_; KartLQRDynamics* cur_;
    std::vector<Vector> otgt_; std::vector<Weights> otw_;
    std::map<int, std::vector<double>> aw_; std::map<int, std::vector<int>> ai_;
    std::vector<KartLQRDynamics*> avoid_;
    Matrix Q_, R_; Vector q_;
};

class KartLQR {
public:
    // Drop-in for KartLQR.cs:17 — returns player 0's first control (2 values), solved on the GPU.
    static Vector solveFeedbackLQR(const std::vector<KartLQRDynamics*>& dynamics, const std::vector<KartLQRCosts*>& costs,
                                   const std::vector<Vector>& initials, int horizon)
    {
        const int N = (int)dynamics.size(), n = 4 * N;
        if (N < 1 || N > HK_MAX_PLAYERS || (int)costs.size() != N || (int)initials.size() != N) throw std::invalid_argument("player count mismatch");
        std::vector<double> A((size_t)N * 16), B((size_t)N * 8), Q((size_t)N * n * n), q((size_t)N * n), R((size_t)N * 4), x0(n), u0(2 * N);
        for (int i = 0; i < N; ++i) {
            if (dynamics[i]->getXDim() != 4 || dynamics[i]->getUDim() != 2) throw std::invalid_argument("only 4-state / 2-control players");
            copy(dynamics[i]->getA(), 4, 4, &A[(size_t)i * 16]);
            copy(dynamics[i]->getB(), 4, 2, &B[(size_t)i * 8]);
            copy(costs[i]->getQMatrix(), n, n, &Q[(size_t)i * n * n]);
            copy(costs[i]->getRMatrix(), 2, 2, &R[(size_t)i * 4]);
            const Vector& qv = costs[i]->getQVec();
            if ((int)qv.size() != n || (int)initials[i].size() != 4) throw std::invalid_argument("dimension mismatch");
            for (int c = 0; c < n; ++c) q[(size_t)i * n + c] = qv[c];
            for (int c = 0; c < 4; ++c) x0[4 * i + c] = initials[i][c];
        }
        hk_check(hk_lqng_solve_one(N, horizon, A.data(), B.data(), Q.data(), q.data(), R.data(), x0.data(), u0.data()));
        return Vector{u0[0], u0[1]};
    }
private:
    static void copy(const Matrix& m, int rows, int cols, double* dst)
    {
        if (m.rows != rows || m.cols != cols) throw std::invalid_argument("dimension mismatch");
        for (size_t e = 0; e < m.a.size(); ++e) dst[e] = m.a[e];
    }
};

}}}  // namespace KartGame::AI::LQR
