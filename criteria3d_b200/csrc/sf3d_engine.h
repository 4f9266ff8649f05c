// sf3d_engine.h -- host-side control of the time step: the Picard / retry / time-step logic of
// CPUSolver::waterMainLoop, waterApproximationLoop, solveLinearSystem, checkCourant
// (agrolib/soilFluxes3D/cpusolver.cpp:143-190, 392-468, 672-703, 248-281) and
// Water::evaluateWaterBalance / acceptStep / restoreBestStep (water.cpp:165-267).
// Only scalars cross the host/device boundary here; all per-node work is in the kernels.
#pragma once
#include "sf3d_backend.h"

namespace sf3d {

struct Balance {            // balanceData_t, types.h:175-184
    double waterStorage = 0., waterSinkSource = 0., waterMBE = 0., waterMBR = 0.;
    double heatStorage = 0., heatSinkSource = 0., heatMBE = 0., heatMBR = 0.;
};

enum class BalanceResult { Accepted, Refused, Halved, Nan };   // balanceResult_t, types.h:174

SolverParams default_params();              // SolverParameters defaults, types.h:291-315

struct Engine {
    SF3DView v{};
    SolverParams *p = nullptr;              // persists across clean/initialize like CPUSolverObject
    Balance curStep, prevStep, curPeriod, wholePeriod;
    double bestMBRerror = SF3D_NODATA;
    double courantWater = 0.;
    sf3d_counters cnt{};
    int xcur = 0;                           // which of x0/x1 holds the current solution vector
    int lastSweeps = 6;                     // sweeps of the previous solve (launch batching hint)
    // the accept pass of the previous step already did the next try's first pass (oldH = H, x = H, SeOld = Se); cleared
    // by anything that touches the state in between (setters, restore, re-initialisation)
    bool tryPrepared = false;
    bool computeWater = true;
    bool computeHeat = false;
    // per-node heat coefficients (kern_heat_coeffs) currently stored for this heat sub-step length, with the current
    // temperatures: the first heatLoop after the flux snapshot does not recompute them
    bool heatCoeffsCurrent = false;
    double heatCoeffsDt = 0.;

    double *xbuf(int k) const { return k ? v.x1 : v.x0; }

    // soilFluxes3D.cpp:1785-1821 / 1760-1777
    double computeStep(double maxTimeStep);
    void   computePeriod(double timePeriod);

    // water
    bool          waterMainLoop(double maxTimeStep, double &acceptedTimeStep);
    BalanceResult waterApproximationLoop(double deltaT);
    int           solveWater(int approx, double deltaT, Ctrl *seen, bool *postDone);   // returns final SOLVE_* status
    bool          courantFailed(double deltaT, double courant);
    BalanceResult evaluateWaterBalance(int approx, double deltaT, const Ctrl *seen = nullptr);
    void          computeCurrentMassBalance(double deltaT, const Ctrl &c);
    void          acceptStep(double deltaT);
    void          restoreBestStep(double deltaT);
    double        totalWaterContent();
    double        totalBoundaryWaterFlow(uint32_t boundaryType);
    uint8_t       initializeWaterBalance();
    void          updateWaterBalanceDataWholePeriod();

    // heat (soilFluxes3D.cpp:1800-1818, cpusolver.cpp:77-91, 471-605, heat.cpp:237-413)
    bool    updateBoundaryHeatData(double maxTimeStep, double &actualTimeStep);
    void    runHeat(double maxTimeStep, double dtWater);
    bool    heatLoop(double timeStepHeat, double timeStepWater);
    uint8_t initializeHeatBalance();
    void    updateHeatBalanceDataWholePeriod();

    uint32_t calcCurrentMaxIterationNumber(int approx) const;   // solver.h:55-59
};

}  // namespace sf3d
