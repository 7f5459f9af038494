"""Probe (run where /root/reference exists): the label / node-id order the reference generates for a Plan_A_Matrix depends on
PYTHONHASHSEED (labels_for_grap ends with list(set(nodes_plan_b)), generate_neo4j_multi_hpf.py:182-183), see DESIGN.md section 7.
    echo "[[1,2,3,4,5],[1,2,3,4],[2,3,4,5],[1,2],[3,4]]" | PYTHONHASHSEED=1 python tests/golden/probe_plan_a_labels.py"""
import sys, json, tempfile
sys.path.insert(0, '/root/reference')
sys.argv = ['x']
from graph_generation.generate_neo4j_multi_hpf import labels_for_grap
conf = {"Plan_A_Matrix": json.loads(sys.stdin.read())}
d = tempfile.mkdtemp() + "/"
allc, a, b, top = labels_for_grap(conf, "12345", d)
print("all", allc, "| plan_b", b, "| top_b", top)
