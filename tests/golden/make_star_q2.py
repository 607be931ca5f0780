"""Writes tests/golden/star_q2.json and star_q3.json from the reference's data/star-q2.mesh and
data/star-q3.mesh (curved quadrilateral meshes, nodes in the legacy `Quadratic` / `Cubic`
collections): element vertex lists, boundary segments and the nodal values.  star-q2 is the input
of the reference's known answer remhos_tests.cpp:88-91, star-q3 the mesh of BASELINE config 5.
The GPU box has no /root/reference; tests materialise the mesh file from this fixture with
`materialise()` below (our own MFEM mesh v1.0 writer).

Run here:  python tests/golden/make_star_q2.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = {'star_q2.json': '/root/reference/data/star-q2.mesh', 'star_q3.json': '/root/reference/data/star-q3.mesh'}


def parse(path):
    toks = []
    with open(path) as f:
        assert f.readline().startswith('MFEM mesh v1.0')
        for line in f:
            toks.extend(line.split('#')[0].split())
    pos = toks.index('elements') + 1
    ne = int(toks[pos]); pos += 1
    elems = []
    for _ in range(ne):
        elems.append([int(t) for t in toks[pos:pos + 6]]); pos += 6
    assert toks[pos] == 'boundary'
    nb = int(toks[pos + 1]); pos += 2
    bdr = []
    for _ in range(nb):
        bdr.append([int(t) for t in toks[pos:pos + 4]]); pos += 4
    assert toks[pos] == 'vertices'
    nv = int(toks[pos + 1]); pos += 2
    assert toks[pos] == 'nodes' and toks[pos + 3] in ('Quadratic', 'Cubic')
    assert toks[pos + 4:pos + 8] == ['VDim:', '2', 'Ordering:', '0']
    vals = toks[pos + 8:]
    return dict(dimension=2, elements=elems, boundary=bdr, vertices=nv, collection=toks[pos + 3],
                vdim=2, ordering=0, nodes=vals)


def materialise(path, fixture=os.path.join(HERE, 'star_q2.json')):
    with open(fixture) as f:
        d = json.load(f)
    with open(path, 'w') as f:
        f.write('MFEM mesh v1.0\n\ndimension\n%d\n\nelements\n%d\n' % (d['dimension'], len(d['elements'])))
        for e in d['elements']:
            f.write(' '.join(str(x) for x in e) + '\n')
        f.write('\nboundary\n%d\n' % len(d['boundary']))
        for b in d['boundary']:
            f.write(' '.join(str(x) for x in b) + '\n')
        f.write('\nvertices\n%d\n\nnodes\nFiniteElementSpace\nFiniteElementCollection: %s\nVDim: %d\n'
                'Ordering: %d\n\n' % (d['vertices'], d['collection'], d['vdim'], d['ordering']))
        f.write('\n'.join(d['nodes']) + '\n')
    return path


if __name__ == '__main__':
    for name, src in SRC.items():
        with open(os.path.join(HERE, name), 'w') as f:
            json.dump(parse(src), f)
        print('wrote', name)
