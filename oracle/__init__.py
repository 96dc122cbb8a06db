"""CPU oracle for the template-switch alignment path -- TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this package (as the checker).  The product
package template_switch_aligner_b200 never does.
"""
