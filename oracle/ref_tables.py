"""Rule tables of the live reference as a JSON fixture (TEST INFRASTRUCTURE; needs /root/reference, build container only).

    python oracle/ref_tables.py      -> tests/golden/ref_tables.json

reward[phase][subnet] = [LWF, ASF, RIA] is read from the reference's own BlueRewardMachine.get_phase_rewards
(CybORG/Shared/BlueRewardMachine.py:17-67) by CALLING it, so the 81 numbers in the fixture are the reference's, not a
transcription; durations are the `duration` attributes of instantiated reference actions.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

H._ensure_path()
from CybORG.Shared.BlueRewardMachine import BlueRewardMachine  # noqa: E402
from CybORG.Simulator.Actions import Analyse, DeployDecoy, Monitor, Remove, Restore, Sleep  # noqa: E402
from CybORG.Simulator.Actions.AbstractActions import (AggressiveServiceDiscovery, DegradeServices, DiscoverDeception, DiscoverRemoteSystems,  # noqa: E402
                                                      ExploitRemoteService, Impact, PrivilegeEscalate, StealthServiceDiscovery)
from CybORG.Simulator.Actions.ConcreteActions.ControlTraffic import AllowTrafficZone, BlockTrafficZone  # noqa: E402

brm = BlueRewardMachine("Blue")
reward = {str(p): {sn: [v["LWF"], v["ASF"], v["RIA"]] for sn, v in brm.get_phase_rewards(p).items()} for p in range(3)}
kw = dict(session=0, agent="a")
import ipaddress
ip, net = ipaddress.IPv4Address("10.0.0.1"), ipaddress.IPv4Network("10.0.0.0/24")
acts = {"Sleep": Sleep(), "Monitor": Monitor(**kw), "Analyse": Analyse(hostname="h", **kw), "Remove": Remove(hostname="h", **kw),
        "Restore": Restore(hostname="h", **kw), "DeployDecoy": DeployDecoy(hostname="h", **kw),
        "BlockTrafficZone": BlockTrafficZone(from_subnet="a", to_subnet="b", **kw), "AllowTrafficZone": AllowTrafficZone(from_subnet="a", to_subnet="b", **kw),
        "DiscoverRemoteSystems": DiscoverRemoteSystems(subnet=net, **kw), "AggressiveServiceDiscovery": AggressiveServiceDiscovery(ip_address=ip, **kw),
        "StealthServiceDiscovery": StealthServiceDiscovery(ip_address=ip, **kw), "DiscoverDeception": DiscoverDeception(ip_address=ip, **kw),
        "ExploitRemoteService": ExploitRemoteService(ip_address=ip, **kw), "PrivilegeEscalate": PrivilegeEscalate(hostname="h", **kw),
        "Impact": Impact(hostname="h", **kw), "DegradeServices": DegradeServices(hostname="h", **kw)}
out = {"reward": reward, "duration": {k: int(a.duration) for k, a in acts.items()}, "cost": {k: float(a.cost) for k, a in acts.items()},
       "priority": {k: int(a.priority) for k, a in acts.items()}}
path = os.path.join(HERE, "..", "tests", "golden", "ref_tables.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(path, {k: len(v) for k, v in out.items()})
