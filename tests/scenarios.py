"""Hand-derived micro-scenarios for the three collision systems.  Every expected outcome follows from the
rule text of SURVEY.md section 8a (and DESIGN.md "collision rules"), not from running any implementation.
Maps use the upstream string syntax ('.' free, '#' obstacle, a-z agents, A-Z their targets); agent index =
alphabetical order.  Actions: 0 stay, 1 up (-1,0), 2 down (+1,0), 3 left (0,-1), 4 right (0,+1).
`expect[system]` lists the (x, y) of every agent after ONE step (unpadded coordinates)."""

SCENARIOS = [
    dict(name="head_on_swap",
         map="""
             ab...
             .....
             ...BA
         """,
         actions=[4, 3],
         expect={s: [(0, 0), (0, 1)] for s in ("priority", "block_both", "soft")}),
    dict(name="follow_lower_index_leads",
         # a leads, b and c follow to the right
         map="""
             cba..
             .....
             ..ABC
         """,
         actions=[4, 4, 4],
         expect={"priority": [(0, 3), (0, 2), (0, 1)],     # each vacated cell is free when the next index moves
                 "soft": [(0, 3), (0, 2), (0, 1)],          # following is allowed
                 "block_both": [(0, 3), (0, 1), (0, 0)]}),  # a cell occupied before the step is blocked
    dict(name="follow_higher_index_leads",
         map="""
             abc..
             .....
             ..ABC
         """,
         actions=[4, 4, 4],
         expect={"priority": [(0, 0), (0, 1), (0, 3)],      # a and b see the not-yet-moved agent in front
                 "soft": [(0, 1), (0, 2), (0, 3)],
                 "block_both": [(0, 0), (0, 1), (0, 3)]}),
    dict(name="two_into_one",
         map="""
             a.b..
             .....
             ...AB
         """,
         actions=[4, 3],
         expect={"priority": [(0, 1), (0, 2)],              # index 0 moves first and takes the cell
                 "soft": [(0, 1), (0, 2)],                  # lowest index keeps the contested cell
                 "block_both": [(0, 0), (0, 2)]}),          # both claimants are blocked
    dict(name="into_obstacle_and_wall",
         map="""
             a#.b.
             .....
             ...AB
         """,
         actions=[4, 1],                                      # a into '#', b up into the border ring
         expect={s: [(0, 0), (0, 3)] for s in ("priority", "block_both", "soft")}),
    dict(name="rotation_cycle",
         map="""
             ab...
             dc...
             ..ABC
             ....D
         """,
         actions=[4, 2, 3, 1],                                # a right, b down, c left, d up around a 2x2 block
         expect={"priority": [(0, 0), (0, 1), (1, 1), (1, 0)],
                 "soft": [(0, 1), (1, 1), (1, 0), (0, 0)],   # no vertex and no edge conflict: everybody moves
                 "block_both": [(0, 0), (0, 1), (1, 1), (1, 0)]}),
    dict(name="cascade_behind_blocked_agent",
         map="""
             ab#..
             .....
             ...AB
         """,
         actions=[4, 4],                                      # b runs into '#', so a cannot take b's cell
         expect={s: [(0, 0), (0, 1)] for s in ("priority", "block_both", "soft")}),
    dict(name="contest_then_cascade",
         map="""
             a.bc.
             .....
             ..ABC
         """,
         actions=[4, 3, 3],                                   # a and b contest (0,1); c follows b
         expect={"priority": [(0, 1), (0, 2), (0, 3)],
                 "soft": [(0, 1), (0, 2), (0, 3)],
                 "block_both": [(0, 0), (0, 2), (0, 3)]}),
]


def clean_map(text):
    return "\n".join(line.strip() for line in text.strip().splitlines())
