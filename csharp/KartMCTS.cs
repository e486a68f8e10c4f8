// KartMCTS.cs — drop-in replacement of Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs.
// Public surface kept verbatim (reference KartMCTS.cs:18-38, 50, 80, 108, 167, 204-236): KartMCTSNode, constructSearchTree (x2),
// getBestStatesSequence, upperConfidenceStrategy, NextGaussian (x3).  `parallel` keeps the reference's meaning:
//   parallel == false (what HierarchicalKartAgent passes, HierarchicalKartAgent.cs:250,271): the sequential search — findLeaf, ONE
//     simulate() whose every state becomes a node (:271-276), backpropagate from its terminal node — runs on the GPU
//     (hk_mcts_forest_search, a device-resident tree that constructSearchTree(root) continues, :80-106) and comes back as the same
//     KartMCTSNode graph the reference would have built (children in insertion order, float totalValue, numEpisodes);
//   parallel == true: the reference's leaf-parallel processLeaf (:124-159) with `RolloutsPerChild` playouts per child in ONE launch,
//     tree in C#.
// The wall-clock budget T becomes an iteration count: T * IterationsPerSecond (a calibration of how fast the C# simulate() was).
using System;
using System.Collections.Generic;
using System.Diagnostics;
using System.Linq;
using UnityEngine;
using MathNet.Numerics.Distributions;
using KartGame.AI.Native;

namespace KartGame.AI.MCTS
{
    public class KartMCTSNode
    {
        public DiscreteGameState state; public KartMCTSNode parent; public Dictionary<DiscreteKartAction, KartMCTSNode> children;
        public float totalValue; public int numEpisodes; public int childrenAsRoot; public string createdBy;
        public KartMCTSNode(DiscreteGameState state, KartMCTSNode parent = null, string createdBy = "")
        {
            this.state = state; this.parent = parent; children = new Dictionary<DiscreteKartAction, KartMCTSNode>();
            totalValue = 0.0f; numEpisodes = 0; childrenAsRoot = 0; this.createdBy = createdBy;
        }
    }

    public class KartMCTS
    {
        static System.Random random = new System.Random();
        static Normal normalDist = new Normal();
        public static long RolloutsPerChild = 4096;
        static readonly Dictionary<RacingEnvController, IntPtr> games = new Dictionary<RacingEnvController, IntPtr>();
        static ulong seedCounter = (ulong)DateTime.Now.Ticks;

        public static double IterationsPerSecond = 600.0;
        public static int ReserveSearches = 3;                            // CyclesRootProcessed < 3 (HierarchicalKartAgent.cs:265)
        sealed class DeviceTree { public IntPtr forest; ~DeviceTree() { if (forest != IntPtr.Zero) HkNative.hk_mcts_forest_destroy(forest); } }
        static readonly System.Runtime.CompilerServices.ConditionalWeakTable<KartMCTSNode, DeviceTree> deviceTrees =
            new System.Runtime.CompilerServices.ConditionalWeakTable<KartMCTSNode, DeviceTree>();

        public static KartMCTSNode constructSearchTree(DiscreteGameState state, double T = 0.09, bool parallel = false)
        {
            return constructSearchTree(new KartMCTSNode(state), T, parallel);
        }

        // parallel == false: reference :61-66 / :91-96 on the device.  The graph under `root` is rebuilt from the device's node records.
        static KartMCTSNode constructSequential(KartMCTSNode root, int iterations)
        {
            var game = GameOf(root.state); var rs = Pack(root.state);
            int plies = Math.Max(1, root.state.kartStates.Sum(k => Math.Max(0, root.state.finalSection - k.section)));
            int[] fresh = { 0 };
            if (!deviceTrees.TryGetValue(root, out var tree))
            {
                tree = new DeviceTree();
                HkNative.Check(HkNative.hk_mcts_forest_create(game, 1, 1 + ReserveSearches * iterations * plies, out tree.forest));
                deviceTrees.Add(root, tree); fresh[0] = 1;
            }
            var best = new HkGameState[HkNative.MaxSeq]; var nBest = new int[1]; var nNodes = new int[1]; var status = new int[1];
            ulong seed; lock (games) { seed = seedCounter++; }
            HkNative.Check(HkNative.hk_mcts_forest_search(tree.forest, new[] { rs }, fresh, iterations, seed, best, nBest, nNodes, status));
            if (status[0] == 2) throw new DivideByZeroException();
            var rec = new HkNative.HkMctsNode[nNodes[0]];
            HkNative.Check(HkNative.hk_mcts_forest_nodes(tree.forest, 0, rec, rec.Length, out int n));
            var parent = new int[n]; var depth = new int[n]; parent[0] = -1; int maxDepth = 0;
            for (int i = 0; i < n; i++)
                for (int c = rec[i].first_child; c >= 0; c = rec[c].next_sibling) { parent[c] = i; depth[c] = depth[i] + 1; maxDepth = Math.Max(maxDepth, depth[c]); }
            int b = root.state.gameParams.velocityBucketSize, vmax = (int)root.state.kartAgents[0].m_Kart.GetMaxSpeed();
            Func<int, DiscreteKartAction> actionOf = gi => new DiscreteKartAction { min_velocity = 6 + (gi >> 2) * b, max_velocity = Math.Min(6 + (gi >> 2) * b + b, vmax), lane = (gi & 3) + 1 };
            root.children.Clear(); root.totalValue = rec[0].totalValue; root.numEpisodes = rec[0].numEpisodes; root.childrenAsRoot = n - 1;
            var node = new KartMCTSNode[n]; node[0] = root;
            for (int i = 0; i < n; i++)                                   // creation order: a parent precedes its children; states by makeMove as in :273
                for (int c = rec[i].first_child; c >= 0; c = rec[c].next_sibling)
                {
                    var a = actionOf(rec[c].gen);
                    var ch = new KartMCTSNode(node[i].state.makeMove(a), node[i]) { totalValue = rec[c].totalValue, numEpisodes = rec[c].numEpisodes };
                    node[i].children[a] = ch; node[c] = ch;               // insertion order = the reference Dictionary's enumeration order
                }
            return root;
        }

        public static KartMCTSNode constructSearchTree(KartMCTSNode root, double T = 0.09, bool parallel = false)
        {
            if (!parallel) return constructSequential(root, Math.Max(1, (int)(T * IterationsPerSecond)));
            var timer = new Stopwatch(); double total = 0.0f;
            while (total < T)
            {
                timer.Reset(); timer.Start();
                processLeaf(findLeaf(root), root);
                timer.Stop(); total += timer.Elapsed.TotalSeconds;
            }
            return root;
        }

        public static List<DiscreteGameState> getBestStatesSequence(KartMCTSNode node)          // reference :108-122, unchanged
        {
            var bestStates = new List<DiscreteGameState>();
            try
            {
                while (node.children.Count > 0)
                {
                    node = node.children[upperConfidenceStrategy(node)];
                    if (node.state.kartStates.All((s) => s.section == node.state.lastCompletedSection)) bestStates.Add(node.state);
                }
            }
            catch (DivideByZeroException) { }
            return bestStates;
        }

        static IntPtr GameOf(DiscreteGameState s)
        {
            lock (games)
            {
                if (games.TryGetValue(s.envController, out var h)) return h;
                var env = s.envController;
                var sec = env.Sections.Select(t => new HkSection { insideR = t.trackInsideRadius, length = t.trackLength, width = t.trackWidth,
                    turnDeg = t.turnDegrees, leftTurn = t.leftTurn ? 1 : 0, optimalLane = t.optimalLane }).ToArray();
                Func<KartAgent, HkKart> kart = a => new HkKart { accel = a.m_Kart.m_FinalStats.Acceleration, braking = a.m_Kart.m_FinalStats.Braking,
                    topSpeed = a.m_Kart.m_FinalStats.TopSpeed, reverseSpeed = a.m_Kart.m_FinalStats.ReverseSpeed, maxGs = a.m_Kart.m_FinalStats.MaxGs,
                    minGs = a.m_Kart.m_FinalStats.MinGs, tireWearFactor = a.m_Kart.m_FinalStats.TireWearFactor };
                var karts = s.kartAgents.Select(kart).ToArray();          // DiscreteGameState.kartAgents order (nextMoves, :326)
                var envKarts = env.Agents.Select(kart).ToArray();         // envController.Agents order (applyAction, :129)
                var p = new HkGameParams { velocityBucketSize = s.gameParams.velocityBucketSize, timePrecision = s.gameParams.timePrecision,
                    sectionWindow = s.gameParams.sectionWindow, treeSearchDepth = s.gameParams.treeSearchDepth, maxLaneChanges = env.MaxLaneChanges,
                    collisionWindow = s.gameParams.collisionWindow, teamScoreRewardMultiplier = env.TeamScoreRewardMultiplier, maxEpisodeSteps = env.maxEpisodeSteps };
                HkNative.Check(HkNative.hk_game_create(sec, sec.Length, karts, karts.Length, envKarts, envKarts.Length, ref p, out h));
                games[env] = h;
                return h;
            }
        }

        static HkGameState Pack(DiscreteGameState s)
        {
            var g = new HkGameState { n_karts = s.kartStates.Count, initialSection = s.initialSection, lastCompletedSection = s.lastCompletedSection, finalSection = s.finalSection };
            for (int i = 0; i < s.kartStates.Count; i++)
            {
                var k = s.kartStates[i];
                g.Set(i, new HkKartState { player = k.player, team = k.team, section = k.section, timeAtSection = k.timeAtSection, min_velocity = k.min_velocity,
                    max_velocity = k.max_velocity, lane = k.lane, tireAge = k.tireAge, laneChanges = k.laneChanges, infeasible = k.infeasible ? 1 : 0 });
            }
            return g;
        }

        // GPU form of processLeaf (:124-159): expand every child, R rollouts from each in one launch, backpropagate the sums.
        static void processLeaf(KartMCTSNode node, KartMCTSNode root)
        {
            var over = node.state.isOver();
            if (over.Item1) { backpropagate(node, over.Item2, 1); return; }
            var moves = node.state.nextMoves();
            var kids = new KartMCTSNode[moves.Count];
            for (int j = 0; j < moves.Count; j++)
            {
                if (!node.children.ContainsKey(moves[j])) { node.children[moves[j]] = new KartMCTSNode(node.state.makeMove(moves[j]), node); root.childrenAsRoot += 1; }
                kids[j] = node.children[moves[j]];
            }
            int n = kids.Length, K = node.state.kartStates.Count;
            var leaves = kids.Select(k => Pack(k.state)).ToArray();
            var visit = new long[n * HkNative.MaxActions]; var nan = new long[n * HkNative.MaxActions]; var plies = new long[n];
            var reward = new double[n * HkNative.MaxActions * HkNative.MaxKarts];
            ulong seed; lock (games) { seed = seedCounter++; }
            HkNative.Check(HkNative.hk_mcts_rollouts_multi(GameOf(node.state), leaves, n, RolloutsPerChild, seed, 0, visit, reward, nan, plies));
            for (int j = 0; j < n; j++)
            {
                long cnt = 0; var sum = new double[HkNative.MaxKarts];
                for (int a = 0; a < HkNative.MaxActions; a++)
                {
                    cnt += visit[j * HkNative.MaxActions + a] - nan[j * HkNative.MaxActions + a];
                    for (int k = 0; k < K; k++) sum[k] += reward[(j * HkNative.MaxActions + a) * HkNative.MaxKarts + k];
                }
                if (cnt == 0) { var o = kids[j].state.isOver(); if (o.Item1) backpropagate(kids[j], o.Item2, (int)RolloutsPerChild); continue; }
                for (var nd = kids[j]; nd != null; nd = nd.parent)       // backpropagate (:280-289) with summed results
                {
                    nd.totalValue += (float)sum[nd.state.upNext()];
                    nd.numEpisodes += (int)cnt;
                }
            }
        }

        static void backpropagate(KartMCTSNode node, List<float> result, int count)
        {
            while (node != null) { node.totalValue += result[node.state.upNext()] * count; node.numEpisodes += count; node = node.parent; }
        }

        private static float UCTWeight(KartMCTSNode node)                                        // reference :162-165, unchanged
        {
            return (node.totalValue / node.numEpisodes) + Mathf.Sqrt(1.0f) * (Mathf.Log(node.parent.numEpisodes / node.numEpisodes));
        }

        public static DiscreteKartAction upperConfidenceStrategy(KartMCTSNode node)              // reference :167-192, unchanged
        {
            int index = random.Next(node.children.Count);
            DiscreteKartAction best = node.children.Keys.ElementAt(index);
            float best_uct = UCTWeight(node.children[best]);
            foreach (var item in node.children)
            {
                float node_uct = UCTWeight(item.Value);
                if (node_uct > best_uct) { best_uct = node_uct; best = item.Key; }
            }
            return best;
        }

        private static KartMCTSNode findLeaf(KartMCTSNode root)                                  // reference :194-201, unchanged
        {
            while (root.children.Count > 0 && root.children.Count == root.state.nextMoves().Count) root = root.children[upperConfidenceStrategy(root)];
            return root;
        }

        public static float NextGaussian()                                                       // reference :204-217
        {
            float v1, v2, s;
            do { v1 = 2.0f * (float)random.NextDouble() - 1.0f; v2 = 2.0f * (float)random.NextDouble() - 1.0f; s = v1 * v1 + v2 * v2; } while (s >= 1.0f || s == 0f);
            s = Mathf.Sqrt((-2.0f * Mathf.Log(s)) / s);
            return v1 * s;
        }
        public static float NextGaussian(float mean, float standard_deviation) { return mean + (float)normalDist.Sample() * standard_deviation; }
        public static float NextGaussian(float mean, float standard_deviation, float min, float max)
        {
            float x; int attempts = 0;
            do { x = NextGaussian(mean, standard_deviation); attempts += 1; } while ((x < min || x > max) && attempts < 10);
            if (attempts == 10 && (x < min || x > max)) return mean;
            return x;
        }
    }
}
