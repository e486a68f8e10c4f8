"""Experiment-log text in the format `experiment_log_parser.py` of the reference reads.

The reference appends `TelemetryViewer`'s text to `ExperimentLogs/<name>.txt` after each race
(Assets/Karting/Scripts/RacingEnvController.cs:249-265,289-305; Assets/Karting/Scripts/TelemetryViewer.cs:50-104).  This module
writes the same lines from the bookkeeping the headless race loop keeps (hk_race_kart / hk_race_plan), so that races run on the
GPU can be scored by the reference's own parser.  Collisions are always 0: the kinematic stand-in has no contacts.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
FIXED_DT = F32(0.02)                       # Time.fixedDeltaTime


def _f(x) -> str:
    """C# float.ToString(): up to 7 significant digits."""
    s = "%.7g" % float(F32(x))
    return s


def lap_times(plan, n_sections: int, laps: int):
    """(last, best) lap time as TelemetryViewer.cs:60-72 derives them from the lap-completion steps."""
    last, best, prev = F32(0), F32(0), 0
    for k in range(min(laps, len(plan["lapStep"]))):
        step = int(plan["lapStep"][k])
        if step <= 0:
            break
        last = FIXED_DT * F32(step - prev)
        if best < 10 or last < best:
            best = last
        prev = step
    return last, best


def race_text(names, karts_race, plans_race, n_sections: int, laps: int, episode_steps: int) -> str:
    """The text block of one race: one group of lines per agent, then the winner line (TelemetryViewer.cs:50-104)."""
    lines, winner, min_time = [], "", F32(1000)
    for name, k, p in zip(names, karts_race, plans_race):
        done_laps = int(k["section"]) // n_sections
        last, best = lap_times(p, n_sections, laps)
        total = FIXED_DT * F32(int(k["sectionStep"]) if not k["active"] else episode_steps)
        if not k["active"] and winner != "Tie":
            if total < min_time:
                winner, min_time = name, total
            elif total == min_time:
                winner = "Tie"
        lines += [f"{name} Speed: {_f(k['v'])}", f"{name} Last Lap: {_f(last)}", f"{name} Best Lap: {_f(best)}",
                  f"{name} Total Time: {_f(total)}", f"{name} Laps Completed: {done_laps}/{laps}",
                  f"{name} Illegal Lane Changes: {int(k['illegalLaneChanges'])}", f"{name} Collisions: 0",
                  f"{name} Avg Target Lane Difference: {_f(p['avgLaneDiff'])}", f"{name} Avg Target Vel Difference: {_f(p['avgVelDiff'])}"]
    lines.append("Winner: " + winner)
    return "\n".join(lines) + "\n"


def write_experiment_log(path: str, names, karts, plans, n_sections: int, laps: int, episode_steps: int, append: bool = False) -> None:
    """One `Experiment k` block per race, as RacingEnvController.cs:261-264 writes them."""
    with open(path, "a" if append else "w") as f:
        for r in range(karts.shape[0]):
            f.write(f"Experiment {r}\n")
            f.write(race_text(names, karts[r], plans[r], n_sections, laps, episode_steps) + "\n")
