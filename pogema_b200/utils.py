"""upstream utils.py :: render_grid - plain text rendering of one instance."""


def render_grid(obstacles, positions_xy=None, targets_xy=None, is_active=None):
    rows = []
    pos = {tuple(p): i for i, p in enumerate(positions_xy or []) if is_active is None or is_active[i]}
    tgt = {tuple(p): i for i, p in enumerate(targets_xy or []) if is_active is None or is_active[i]}
    for x in range(obstacles.shape[0]):
        line = []
        for y in range(obstacles.shape[1]):
            if (x, y) in pos:
                line.append(chr(ord('a') + pos[(x, y)] % 26))
            elif (x, y) in tgt:
                line.append(chr(ord('A') + tgt[(x, y)] % 26))
            elif obstacles[x, y]:
                line.append('#')
            else:
                line.append('.')
        rows.append(''.join(line))
    return '\n'.join(rows)
