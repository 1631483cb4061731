import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d['e2e']
        print(f, 'registered %.3e'%e['value'], {k:'%.3e'%v['value'] for k,v in e['variants'].items()})
    except Exception as ex:
        print(f, 'ERR', ex)
