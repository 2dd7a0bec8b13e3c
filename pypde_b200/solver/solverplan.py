"""SolverPlan container (pypde/solver/solverplan.py:4-100): ordered lists of rhs / old /
lhs plans executed sequentially."""
from .plans import MetaPlan, PlanRHS


class SolverPlan:
    def __init__(self):
        self.plan_for_lhs = []
        self.plan_for_rhs = []
        self.plan_for_old = []

    def add_rhs(self, plan):
        assert isinstance(plan, PlanRHS)
        self.plan_for_rhs.append(plan)

    def add_old(self, plan):
        assert isinstance(plan, PlanRHS)
        self.plan_for_old.append(plan)

    def add_lhs(self, plan):
        assert isinstance(plan, MetaPlan)
        self.plan_for_lhs.append(plan)

    def solve_rhs(self, b):
        for plan in self.plan_for_rhs:
            b = plan.solve(b)
        return b

    def solve_old(self, b):
        for plan in self.plan_for_old:
            b = plan.solve(b)
        return b

    def solve_lhs(self, b):
        for plan in self.plan_for_lhs:
            b = plan.solve(b)
        return b

    def show_plan(self):
        for title, plans in (("Plans RHS:", self.plan_for_rhs), ("Plans RHS (#2):", self.plan_for_old),
                             ("Plans LHS:", self.plan_for_lhs)):
            if plans:
                print(title)
                for i, p in enumerate(plans):
                    print(i + 1, ")", "Apply method '{:s}' along axis {:1d} ".format(
                        p.flags["method"], p.flags["axis"]))
                print("")
