// sf3d_engine.cpp -- host control loop of the water time step (see sf3d_engine.h).
// The decision tree is the reference's; the per-node work is launched as kernels and only the
// reduced scalars (Courant max, residual status, storage, sink sum) are read back: two to three
// stream synchronisations per Picard approximation.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "sf3d_engine.h"
#include <nvtx3/nvToolsExt.h>

// NVTX ranges per phase of the step (tracing): visible in Nsight Systems / Compute timelines.  Header-only
// NVTX3: without an attached tool the calls are no-ops; SF3D_NVTX=0 removes even those.
namespace {
const bool g_nvtx = !(getenv("SF3D_NVTX") && atoi(getenv("SF3D_NVTX")) == 0);
struct Range {
    explicit Range(const char *name) { if (g_nvtx) nvtxRangePushA(name); }
    ~Range() { if (g_nvtx) nvtxRangePop(); }
};
}

namespace sf3d {

SolverParams default_params()
{
    SolverParams p{};
    p.MBRThreshold = 1e-3;
    p.residualTolerance = 1e-10;
    p.deltaTmin = 1.;
    p.deltaTmax = 600.;
    p.deltaTcurr = SF3D_NODATA;
    p.maxApproximationsNumber = 10;
    p.maxIterationsNumber = 200;
    p.wrcModel = 1;                 // ModifiedVanGenuchten
    p.meanType = 2;                 // Logarithmic
    p.lateralVerticalRatio = 4.;
    p.heatWeightFactor = 0.5;
    p.CourantWaterThreshold = 0.5;
    p.instabilityFactor = 10.;
    return p;
}

// Solver::calcCurrentMaxIterationNumber (solver.h:55-59): float arithmetic, then max(.., 25)
uint32_t Engine::calcCurrentMaxIterationNumber(int approx) const
{
    uint32_t n = static_cast<uint32_t>((approx + 1) * (static_cast<float>(p->maxIterationsNumber)
                                                       / static_cast<float>(p->maxApproximationsNumber)));
    return std::max(n, 25u);
}

// soilFluxes3D.cpp:1785-1821
double Engine::computeStep(double maxTimeStep)
{
    Range r("sf3d computeStep");
    if (computeHeat)
    {
        k_reset_water_fluxes(v);            // resetFluxValues(false, true)
        k_update_conductance(v);            // updateConductance()
    }
    double dtWater;
    if (computeWater)
    {
        // the reference ignores run()'s error code (soilFluxes3D.cpp:1796)
        waterMainLoop(maxTimeStep, dtWater);
    }
    else
        dtWater = std::min(maxTimeStep, p->deltaTmax);

    if (computeHeat)
    {
        double dtHeat = dtWater;
        k_save_water_fluxes(v, dtHeat, dtWater);            // stores the per-node coefficients for (dtHeat, T) first
        heatCoeffsCurrent = true; heatCoeffsDt = dtHeat;
        double dtHeatSum = 0.;
        while (dtHeatSum < dtWater)
        {
            dtHeat = std::min(dtHeat, dtWater - dtHeatSum);
            double reducedTimeStep;
            while (!updateBoundaryHeatData(dtHeat, reducedTimeStep)) dtHeat = reducedTimeStep;
            runHeat(dtHeat, dtWater);
            dtHeatSum += dtHeat;
        }
    }
    ++cnt.steps;
    return dtWater;
}

// Heat::updateBoundaryHeatData, the time-step decision (heat.cpp:322-339)
bool Engine::updateBoundaryHeatData(double maxTimeStep, double &actualTimeStep)
{
    k_boundary_heat(v, maxTimeStep);
    Ctrl c{};
    read_ctrl(v, &c);
    const double courant = c.heatCourantMax;
    const double minTimeStep = p->deltaTmin;
    if (courant > 1. && maxTimeStep > minTimeStep)
    {
        actualTimeStep = std::max(minTimeStep, maxTimeStep / courant);
        if (actualTimeStep > 1.) actualTimeStep = floor(actualTimeStep);
        return false;
    }
    return true;
}

// CPUSolver::run, processType::Heat (cpusolver.cpp:77-91)
void Engine::runHeat(double maxTimeStep, double dtWater)
{
    double dtHeat = maxTimeStep;
    double sumHeatTime = 0;
    while (sumHeatTime < maxTimeStep)
    {
        dtHeat = std::min(dtHeat, maxTimeStep - sumHeatTime);
        if (!heatLoop(dtHeat, dtWater)) dtHeat *= 0.5;
        else sumHeatTime += dtHeat;
    }
}

// CPUSolver::heatLoop (cpusolver.cpp:471-605).  The linear solve is a Jacobi iteration with the
// reference's infinity-norm stopping rule (the reference sweeps Gauss-Seidel sequentially,
// heat.cpp:664-685: same fixed point); it may use up to 4x the reference's sweep
// cap so that it reaches the tolerance whenever Gauss-Seidel does.
bool Engine::heatLoop(double timeStepHeat, double timeStepWater)
{
    Range r("sf3d heat sub-step");
    k_heat_begin(v, timeStepHeat, timeStepWater, heatCoeffsCurrent && heatCoeffsDt == timeStepHeat);   // reset heat fluxes ; x = T ; oldT = T ; C
    heatCoeffsCurrent = false;                          // the solve below changes T
    k_heat_assemble(v, timeStepHeat, timeStepWater);

    const int refCap = (int)calcCurrentMaxIterationNumber((int)p->maxApproximationsNumber - 1);
    const int maxIter = 4 * refCap;
    int launched = 0;
    Ctrl c{};
    int batch = 8;
    for (;;)
    {
        const int n = std::min(batch, maxIter - launched);
        for (int k = 0; k < n; ++k, ++launched)
            k_heat_jacobi(v, xbuf(launched & 1), xbuf((launched + 1) & 1), maxIter, p->residualTolerance);
        read_ctrl(v, &c);
        if (c.status != SOLVE_RUNNING || launched >= maxIter) break;
        batch = std::min(batch * 2, 64);
    }
    cnt.heat_sweeps += (uint64_t)c.sweeps;
    if (c.status != SOLVE_CONVERGED) ++cnt.heat_cap_hits;       // stopped at the cap: the result depends on the sweep (Q6)
    const double *x = xbuf(c.sweeps & 1);

    k_heat_post(v, x, timeStepHeat, timeStepWater, 0);  // T = x ; heat storage ; heat sink sum
    read_ctrl(v, &c);
    // Heat::evaluateHeatBalance (heat.cpp:373-390)
    curStep.heatSinkSource = c.heatSinkSum;
    curStep.heatStorage = c.heatStorage;
    const double deltaHeatStorage = curStep.heatStorage - prevStep.heatStorage;
    curStep.heatMBE = deltaHeatStorage - curStep.heatSinkSource;
    const double referenceHeat = std::max(c.heatStorage * 1e-6, fabs(curStep.heatSinkSource));
    curStep.heatMBR = curStep.heatMBE / referenceHeat;

    if (fabs(curStep.heatMBR) > 1.0 && timeStepHeat > (p->deltaTmin * 10.0))
    {
        k_heat_copy_T(v, 1);                            // restore old temperatures
        return false;
    }
    // Heat::updateHeatBalanceData (heat.cpp:393-398)
    prevStep.heatStorage = curStep.heatStorage;
    prevStep.heatSinkSource = curStep.heatSinkSource;
    curPeriod.heatSinkSource += curStep.heatSinkSource;
    k_heat_accept(v, timeStepHeat, timeStepWater);      // saveHeatFluxValues
    k_heat_copy_T(v, 0);                                // oldT = T
    ++cnt.heat_steps;
    return true;
}

// Heat::initializeHeatBalance (heat.cpp:31-53)
uint8_t Engine::initializeHeatBalance()
{
    wholePeriod.heatSinkSource = curPeriod.heatSinkSource = curStep.heatSinkSource = prevStep.heatSinkSource = 0.;
    wholePeriod.heatMBE = curPeriod.heatMBE = curStep.heatMBE = 0.;
    k_heat_post(v, nullptr, 1., 1., 2);
    Ctrl c{};
    read_ctrl(v, &c);
    wholePeriod.heatStorage = curPeriod.heatStorage = curStep.heatStorage = prevStep.heatStorage = c.heatStorage;
    return SF3D_OK;
}

// Heat::updateHeatBalanceDataWholePeriod (heat.cpp:400-413)
void Engine::updateHeatBalanceDataWholePeriod()
{
    wholePeriod.heatSinkSource += curPeriod.heatSinkSource;
    const double deltaStoragePeriod = curStep.heatStorage - curPeriod.heatStorage;
    const double deltaStorageHistorical = curStep.heatStorage - wholePeriod.heatStorage;
    curPeriod.heatMBE = deltaStoragePeriod - curPeriod.heatSinkSource;
    wholePeriod.heatMBE = deltaStorageHistorical - wholePeriod.heatSinkSource;
    const double referenceHeat = std::max(1., fabs(wholePeriod.heatSinkSource));
    wholePeriod.heatMBR = wholePeriod.heatMBE / referenceHeat;
    curPeriod.heatStorage = curStep.heatStorage;
}

// soilFluxes3D.cpp:1760-1777
void Engine::computePeriod(double timePeriod)
{
    double sumCurrentTime = 0.;
    curPeriod.waterSinkSource = 0.;
    curPeriod.heatSinkSource = 0.;
    while (sumCurrentTime < timePeriod)
        sumCurrentTime += computeStep(timePeriod - sumCurrentTime);
    if (computeWater) updateWaterBalanceDataWholePeriod();
    if (computeHeat) updateHeatBalanceDataWholePeriod();
}

// CPUSolver::waterMainLoop (cpusolver.cpp:143-190)
bool Engine::waterMainLoop(double maxTimeStep, double &acceptedTimeStep)
{
    BalanceResult stepStatus = BalanceResult::Refused;
    while (stepStatus != BalanceResult::Accepted)
    {
        Range r("sf3d water try");
        acceptedTimeStep = std::min(p->deltaTcurr, maxTimeStep);
        if (!tryPrepared) k_begin_try(v);                 // oldH = H ; x = H ; Se ; surface capacity
        tryPrepared = false;
        xcur = 0;
        ++cnt.tries;
        if (computeHeat) k_update_conductance(v);       // cpusolver.cpp:171-173

        stepStatus = waterApproximationLoop(acceptedTimeStep);

        if (stepStatus == BalanceResult::Nan) return false;
        if (stepStatus != BalanceResult::Accepted) k_restore_old(v);
    }
    return true;
}

// CPUSolver::checkCourant, the time-step update after a failed test (cpusolver.cpp:262-278)
bool Engine::courantFailed(double /*deltaT*/, double courant)
{
    p->deltaTcurr /= courant;
    int multiply = 0;
    while (p->deltaTcurr < 1.) { p->deltaTcurr *= 10.; ++multiply; }
    p->deltaTcurr = floor(p->deltaTcurr);
    for (int i = 0; i < multiply; i++) p->deltaTcurr /= 10.;
    p->deltaTcurr = std::max(p->deltaTmin, p->deltaTcurr);
    return true;
}

// CPUSolver::solveLinearSystem (cpusolver.cpp:672-703): the stopping rule runs on the device;
// sweeps are enqueued in batches and the control block is read back once per batch.
int Engine::solveWater(int approx, double deltaT, Ctrl *seen, bool *postDone)
{
    Range r("sf3d water solve (Jacobi sweeps)");
    const int maxIter = (int)calcCurrentMaxIterationNumber(approx);
    const int start = xcur;
    int launched = 0;
    Ctrl c{};
    int batch = std::min(std::max(lastSweeps + 1, 4), 32);
    // the post pass rides behind the sweeps and decides on the device whether it runs (kern_post mode 3): the control
    // block read that ends the solve then also carries the balance sums
    const bool follow = k_post_can_follow_solve(v);
    if (k_jacobi_persistent_ok(v))
    {
        // small graph: the whole solve is one launch and one control-block read
        k_jacobi_persistent(v, xbuf(start), xbuf(start ^ 1), maxIter, p->residualTolerance);
        if (follow) k_post_follow_solve(v, start, deltaT, p->deltaTmin);
        read_ctrl(v, &c);
    }
    else for (;;)
    {
        const int n = std::min(batch, maxIter - launched);
        for (int k = 0; k < n; ++k, ++launched)
            k_jacobi(v, xbuf((start + launched) & 1), xbuf((start + launched + 1) & 1), maxIter, p->residualTolerance);
        if (follow) k_post_follow_solve(v, start, deltaT, p->deltaTmin);
        read_ctrl(v, &c);
        if (c.status != SOLVE_RUNNING || launched >= maxIter) break;
        batch = std::min(batch * 2, 32);
    }
    const bool goesOn = c.status == SOLVE_CONVERGED || c.status == SOLVE_MAXITER
                        || (c.status == SOLVE_DIVERGED && !(deltaT > p->deltaTmin));
    *postDone = follow && goesOn;
    *seen = c;
    if (c.status != SOLVE_COURANT_FAIL)
    {
        xcur = (start + c.sweeps) & 1;
        cnt.sweeps += (uint64_t)c.sweeps;
        lastSweeps = c.sweeps;
    }
    courantWater = c.courantMax;
    return c.status;
}

// CPUSolver::waterApproximationLoop (cpusolver.cpp:392-468)
BalanceResult Engine::waterApproximationLoop(double deltaT)
{
    BalanceResult balanceResult = BalanceResult::Refused;
    bestMBRerror = SF3D_NODATA;

    for (int approxIdx = 0; approxIdx < (int)p->maxApproximationsNumber; ++approxIdx)
    {
        ++cnt.approximations;
        Range r("sf3d water approximation");
        k_node_phase(v, deltaT, 1);                             // computeCapacity + updateBoundaryWaterData
        k_assemble(v, deltaT, approxIdx, p->deltaTmin);         // rows + Courant + normalisation
        Ctrl seen{};
        bool postDone = false;
        const int status = solveWater(approxIdx, deltaT, &seen, &postDone);

        if (status == SOLVE_COURANT_FAIL)
        {
            courantFailed(deltaT, courantWater);
            return BalanceResult::Halved;
        }
        const bool isStepValid = (status != SOLVE_DIVERGED);
        if (!isStepValid && deltaT > p->deltaTmin)
        {
            p->deltaTcurr = std::max(p->deltaTmin, p->deltaTcurr / 2.);
            return BalanceResult::Halved;
        }

        if (!postDone) k_post(v, xbuf(xcur), deltaT, 0);        // H = x ; Se ; storage ; sink sum
        balanceResult = evaluateWaterBalance(approxIdx, deltaT, postDone ? &seen : nullptr);

        if (balanceResult == BalanceResult::Accepted || balanceResult == BalanceResult::Halved
            || balanceResult == BalanceResult::Nan)
            return balanceResult;
    }
    return balanceResult;
}

// Water::computeCurrentMassBalance (water.cpp:96-123) from the two reduced sums
void Engine::computeCurrentMassBalance(double deltaT, const Ctrl &c)
{
    Balance cur;
    cur.waterStorage = c.storage;
    const double deltaStorage = cur.waterStorage - prevStep.waterStorage;
    cur.waterSinkSource = c.sinkSum;
    cur.waterMBE = deltaStorage - cur.waterSinkSource;

    const double timePercentage = 0.001 * std::max(deltaT, 30.0) / 3600.;
    double minRefWaterStorage = cur.waterStorage * timePercentage;
    minRefWaterStorage = std::max(minRefWaterStorage, 0.001);
    const double referenceWater = std::max(fabs(cur.waterSinkSource), minRefWaterStorage);
    cur.waterMBR = cur.waterMBE / referenceWater;

    // balanceDataCurrentTimeStep = currentBalance copies the whole struct, heat fields included
    // (default-initialised to 0 in the reference's local)
    curStep = cur;
}

// Water::evaluateWaterBalance (water.cpp:165-227)
BalanceResult Engine::evaluateWaterBalance(int approxNr, double deltaT, const Ctrl *seen)
{
    Ctrl c{};
    if (seen) c = *seen;                    // the post pass followed the solve: its sums came with the solve's control read
    else read_ctrl(v, &c);
    computeCurrentMassBalance(deltaT, c);

    const double currMBRerror = fabs(curStep.waterMBR);

    if (std::isnan(currMBRerror))
    {
        if (deltaT > p->deltaTmin)
        {
            p->deltaTcurr = std::max(p->deltaTcurr * 0.5, p->deltaTmin);
            return BalanceResult::Halved;
        }
        else if (approxNr > 0)
        {
            restoreBestStep(deltaT);
            acceptStep(deltaT);
            return BalanceResult::Accepted;
        }
        return BalanceResult::Nan;
    }

    if (currMBRerror < p->MBRThreshold)
    {
        acceptStep(deltaT);
        if (approxNr < 3 && currMBRerror < p->MBRThreshold * 0.1 && courantWater < p->CourantWaterThreshold)
            p->deltaTcurr = std::min(p->deltaTmax, p->deltaTcurr * 2);
        return BalanceResult::Accepted;
    }

    if (approxNr == 0 || currMBRerror < bestMBRerror)
    {
        dev_copy(v.bestH, v.H, (size_t)v.N * sizeof(double));
        bestMBRerror = currMBRerror;
    }

    if (currMBRerror > (bestMBRerror * p->instabilityFactor) || approxNr == ((int)p->maxApproximationsNumber - 1))
    {
        if (deltaT > p->deltaTmin)
        {
            p->deltaTcurr = std::max(p->deltaTcurr * 0.5, p->deltaTmin);
            return BalanceResult::Halved;
        }
        restoreBestStep(deltaT);
        acceptStep(deltaT);
        return BalanceResult::Accepted;
    }
    return BalanceResult::Refused;
}

// Water::acceptStep (water.cpp:230-251)
void Engine::acceptStep(double deltaT)
{
    prevStep.waterStorage = curStep.waterStorage;
    prevStep.waterSinkSource = curStep.waterSinkSource;
    curPeriod.waterSinkSource += curStep.waterSinkSource;
    static const bool fuse = getenv("SF3D_NO_PREPARED_TRY") == nullptr;
    tryPrepared = fuse && !computeHeat;
    k_accept(v, deltaT, tryPrepared);
}

// Water::restoreBestStep (water.cpp:253-267)
void Engine::restoreBestStep(double deltaT)
{
    k_restore_best(v);              // H = best ; Se
    k_node_phase(v, deltaT, 0);     // K ; updateBoundaryWaterData
    k_post(v, nullptr, deltaT, 2);  // storage ; sink sum
    Ctrl c{};
    read_ctrl(v, &c);
    computeCurrentMassBalance(deltaT, c);
}

// Water::computeTotalWaterContent (water.cpp:71-90)
double Engine::totalWaterContent()
{
    k_post(v, nullptr, 1., 2);
    Ctrl c{};
    read_ctrl(v, &c);
    return c.storage;
}

double Engine::totalBoundaryWaterFlow(uint32_t boundaryType)
{
    k_total_boundary_flow(v, boundaryType);
    Ctrl c{};
    read_ctrl(v, &c);
    return c.boundarySum;
}

// Water::initializeWaterBalance (water.cpp:35-65)
uint8_t Engine::initializeWaterBalance()
{
    const double currentWC = totalWaterContent();
    wholePeriod.waterStorage = currentWC;
    curPeriod.waterStorage = currentWC;
    curStep.waterStorage = currentWC;
    prevStep.waterStorage = currentWC;
    curStep.waterSinkSource = prevStep.waterSinkSource = curPeriod.waterSinkSource = wholePeriod.waterSinkSource = 0.;
    curStep.waterMBR = wholePeriod.waterMBR = 0.;
    curStep.waterMBE = wholePeriod.waterMBE = 0.;
    dev_zero(v.lflow, (size_t)SF3D_NLINK * v.N * sizeof(double));
    dev_zero(v.bSum, (size_t)v.N * sizeof(double));
    return SF3D_OK;
}

// Water::updateWaterBalanceDataWholePeriod (water.cpp:143-156)
void Engine::updateWaterBalanceDataWholePeriod()
{
    wholePeriod.waterSinkSource += curPeriod.waterSinkSource;
    const double deltaStoragePeriod = curStep.waterStorage - curPeriod.waterStorage;
    const double deltaStorageHistorical = curStep.waterStorage - wholePeriod.waterStorage;
    curPeriod.waterMBE = deltaStoragePeriod - curPeriod.waterSinkSource;
    wholePeriod.waterMBE = deltaStorageHistorical - wholePeriod.waterSinkSource;
    const double referenceWater = std::max(0.001, wholePeriod.waterSinkSource);
    wholePeriod.waterMBR = wholePeriod.waterMBE / referenceWater;
    curPeriod.waterStorage = curStep.waterStorage;
}

}  // namespace sf3d
