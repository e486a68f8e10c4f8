// HkNative.cs — P/Invoke bindings of libhk_b200 (include/hk_abi.h). Drop next to the shims in Assets/Karting/Scripts/AI/.
// Blittable structs mirror the C structs field for field ([StructLayout(LayoutKind.Sequential)]); every call returns an
// hk_status (0 = OK) and HkNative.Check turns failures into exceptions carrying hk_last_error().
using System;
using System.Runtime.InteropServices;

namespace KartGame.AI.Native
{
    [StructLayout(LayoutKind.Sequential)] public struct HkSection { public float insideR, length, width, turnDeg; public int leftTurn, optimalLane; }
    [StructLayout(LayoutKind.Sequential)] public struct HkKart { public float accel, braking, topSpeed, reverseSpeed, maxGs, minGs, tireWearFactor; }
    [StructLayout(LayoutKind.Sequential)] public struct HkGameParams
    {
        public int velocityBucketSize, timePrecision, sectionWindow, treeSearchDepth, maxLaneChanges;
        public float collisionWindow, teamScoreRewardMultiplier; public int maxEpisodeSteps;
    }
    [StructLayout(LayoutKind.Sequential)] public struct HkKartState
    {
        public int player, team, section, timeAtSection, min_velocity, max_velocity, lane, tireAge, laneChanges, infeasible;
    }
    [StructLayout(LayoutKind.Sequential)] public struct HkAction { public int min_velocity, max_velocity, lane; }
    [StructLayout(LayoutKind.Sequential)] public struct HkGameState
    {
        public int n_karts, initialSection, lastCompletedSection, finalSection;
        public HkKartState k0, k1, k2, k3;                      // karts[HK_MAX_KARTS]
        public void Set(int i, HkKartState s) { if (i == 0) k0 = s; else if (i == 1) k1 = s; else if (i == 2) k2 = s; else k3 = s; }
    }

    [StructLayout(LayoutKind.Sequential)] public struct HkRaceKart
    {
        public double x, z, v, h; public float steer;
        public int section, lane, laneChanges, illegalLaneChanges, sectionStep, active, team;
    }
    [StructLayout(LayoutKind.Sequential)] public unsafe struct HkRacePlan
    {
        public fixed sbyte lane[64]; public fixed float vel[64]; public fixed sbyte oppLane[64]; public fixed float oppVel[64];
        public fixed int sectionTimes[64]; public fixed int lapStep[8]; public float avgLaneDiff, avgVelDiff;
    }
    [StructLayout(LayoutKind.Sequential)] public struct HkRaceParams
    {
        public double dt; public float accel, braking, coastingDrag, topSpeed, gateHalfWidth;
        public int maxLaneChanges, goalSection, highModeMcts, velocityBucketSize, treeSearchDepth, planEvery, horizon;
    }

    /// <summary>One 2-kart LQNG problem as the reference's providers are constructed (LinearizedBicycle(dt, initial) and
    /// LQRCheckpointReachAvoidCost(target, weights, ...), KartLQRDynamics.cs:22-38, KartLQRCosts.cs:22-55), player 0 = ego: 44 doubles.</summary>
    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct HkLqngRecord2
    {
        public fixed double x0[8];        // [player][x, z, v, h]
        public fixed double target[8];    // [player][x, z, v, h]
        public fixed double tw[8];        // target weights
        public fixed double cw[2];        // control weight
        public fixed double aw[4];        // [player][avoid weight x, z] about the other player
        public fixed double otgt[8];      // [player][the other player's target x, z, v, h]
        public fixed double otw[6];       // [player][weights x, z, v on it]
    }

    public static class HkNative
    {
        const string Lib = "hk_b200";                            // libhk_b200.so / hk_b200.dll on the plugin search path
        public const int MaxActions = 36, MaxKarts = 4, MaxSeq = 16;

        [DllImport(Lib)] public static extern int hk_abi_version();
        [DllImport(Lib)] public static extern int hk_init(int device);
        [DllImport(Lib)] public static extern void hk_shutdown();
        [DllImport(Lib)] static extern IntPtr hk_last_error();
        [DllImport(Lib)] public static extern int hk_lqng_solve_one(int nPlayers, int horizon, double[] A, double[] B, double[] Q, double[] q,
                                                                    double[] R, double[] x0, [Out] double[] u0);
        [DllImport(Lib)] public static extern int hk_lqng_solve_batch(int batch, int nPlayers, int horizon, int timeVarying, double[] A, double[] B,
                                                                      double[] Q, double[] q, double[] R, double[] x0, [Out] double[] u0,
                                                                      [Out] double[] P, [Out] double[] alpha, [Out] double[] traj, [Out] int[] status);
        [DllImport(Lib)] public static extern int hk_lqng_assemble_solve_batch(int batch, int nPlayers, int horizon, double dt, double[] x0, double[] target,
                                                                               double[] tw, double[] cw, double[] aw, double[] otgt, double[] otw,
                                                                               [Out] double[] u0, [Out] int[] status);
        // the same with one blittable 352-byte record per 2-kart problem (x0 | target | tw | cw | aw | otgt | otw, include/hk_abi.h): an array of
        // HkLqngRecord2 is what a C# caller fills naturally, and it crosses PCIe as one copy per chunk
        [DllImport(Lib)] public static extern int hk_lqng_assemble_solve_packed(int batch, int nPlayers, int horizon, double dt, HkLqngRecord2[] records,
                                                                                [Out] double[] u0, [Out] int[] status);
        [DllImport(Lib)] public static extern int hk_game_create(HkSection[] sections, int nSections, HkKart[] karts, int nKarts, HkKart[] envKarts,
                                                                 int nEnvKarts, ref HkGameParams p, out IntPtr game);
        [DllImport(Lib)] public static extern void hk_game_destroy(IntPtr game);
        [DllImport(Lib)] public static extern int hk_mcts_rollouts_multi(IntPtr game, HkGameState[] leaves, int nLeaves, long rolloutsPerLeaf, ulong seed,
                                                                         ulong rolloutOffset, [Out] long[] visit, [Out] double[] rewardSum,
                                                                         [Out] long[] nanCount, [Out] long[] pliesSum);
        // batched tree search (constructSearchTree + getBestStatesSequence, one thread block per root); bestStates [nRoots][16]
        [DllImport(Lib)] public static extern int hk_mcts_search_batch(IntPtr game, HkGameState[] roots, int nRoots, int iterations, int rolloutsPerLeaf,
            ulong seed, [Out] HkGameState[] bestStates, [Out] int[] nBest, [Out] int[] rootEpisodes, [Out] double[] rootValues, [Out] int[] nNodes);

        // The sequential search the reference's callers run (constructSearchTree with parallel == false), device-resident trees that
        // survive between calls like HierarchicalKartAgent.currentRoot; one GPU thread per tree.
        [StructLayout(LayoutKind.Sequential)] public struct HkMctsNode
        { public ulong child_mask; public float totalValue; public int numEpisodes, first_child, last_child, next_sibling; public byte gen, n_legal; public sbyte upnext; public byte pad_; }
        [DllImport(Lib)] public static extern int hk_mcts_forest_create(IntPtr game, int nTrees, int maxNodesPerTree, out IntPtr forest);
        [DllImport(Lib)] public static extern void hk_mcts_forest_destroy(IntPtr forest);
        [DllImport(Lib)] public static extern int hk_mcts_forest_search(IntPtr forest, HkGameState[] roots, int[] fresh, int iterations, ulong seed,
            [Out] HkGameState[] bestStates, [Out] int[] nBest, [Out] int[] nNodes, [Out] int[] status);
        [DllImport(Lib)] public static extern int hk_mcts_forest_nodes(IntPtr forest, int tree, [Out] HkMctsNode[] nodes, int maxNodes, out int nNodes);
        [DllImport(Lib)] public static extern int hk_game_replay_batch(IntPtr game, int batch, int len, HkGameState[] roots, HkAction[] actions,
            [Out] HkGameState[] statesOut, IntPtr upnext, IntPtr over, IntPtr nScores, IntPtr scores, IntPtr nMoves, IntPtr moves, IntPtr movesIndex);

        // the headless loop with the MCTS high level on the device (root states, tree search, waypoint hand-off between two steps)
        [DllImport(Lib)] public static extern int hk_race_run_mcts(IntPtr track, ref HkRaceParams p, IntPtr game, int iterations, int rolloutsPerLeaf, ulong seed,
            int nRaces, int firstStep, int nSteps, [In, Out] HkRaceKart[] karts, [In, Out] HkRacePlan[] plans, [Out] double[] uLast, out long lqngStatusNonzero);
        // headless batch races (kinematic plant instead of PhysX)
        [DllImport(Lib)] public static extern int hk_track_create(HkSection[] sections, double[] triggerXz, double[] forwardXz, double[] laneXz,
                                                                  int nSections, out IntPtr track);
        [DllImport(Lib)] public static extern void hk_track_destroy(IntPtr track);
        [DllImport(Lib)] public static extern int hk_race_recipe(IntPtr track, ref HkRaceParams p, int nRaces, HkRaceKart[] karts, HkRacePlan[] plans,
                                                                 [Out] double[] x0, [Out] double[] target, [Out] double[] tw, [Out] double[] cw,
                                                                 [Out] double[] aw, [Out] double[] otgt, [Out] double[] otw);
        [DllImport(Lib)] public static extern int hk_race_plan_fixed(IntPtr track, ref HkRaceParams p, int nKarts, HkRaceKart[] karts, [In, Out] HkRacePlan[] plans);
        [DllImport(Lib)] public static extern int hk_race_step(IntPtr track, ref HkRaceParams p, int nKarts, int episodeStep, double[] u,
                                                               [In, Out] HkRaceKart[] karts, [In, Out] HkRacePlan[] plans);
        [DllImport(Lib)] public static extern int hk_race_run(IntPtr track, ref HkRaceParams p, int nRaces, int firstStep, int nSteps,
                                                              [In, Out] HkRaceKart[] karts, [In, Out] HkRacePlan[] plans, [Out] double[] uLast,
                                                              out long lqngStatusNonzero);

        public static void Check(int status)
        {
            if (status == 0) return;
            string msg = Marshal.PtrToStringAnsi(hk_last_error());
            if (status == -1) throw new ArgumentException(msg);           // what MathNet throws on a dimension mismatch
            if (status == -5) throw new ArgumentOutOfRangeException(msg); // upNext() == -1, KartDiscreteGame.cs:326
            throw new InvalidOperationException("hk_b200 status " + status + ": " + msg);
        }
    }
}
