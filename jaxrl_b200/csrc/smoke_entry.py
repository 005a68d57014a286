import __graft_entry__ as g
g.smoke()
print("smoke ok")
